"""vadc_b200 -- B200-native Silero VAD v3.1 (16 kHz) engine behind vadc's backend interface.

The product is vadc_b200/libsilero_b200.so (C ABI: include/silero_b200.h, include/vadc_segmenter.h),
built from vadc_b200/csrc/ for sm_100a. This package only binds it for tests and bench.py.
"""
from .api import (CHUNK, SAMPLE_RATE, Engine, Group, EngineError, StreamSegmenter, lib, pinned_empty, pinned_free, seg_params,
                  segments_text, synth_pcm, LIB_PATH, WEIGHTS_PATH, STFT_AUTO, STFT_HYBRID, STFT_EXACT, LSTM_AUTO, LSTM_FP32, LSTM_TENSOR, LAYERS_AUTO, LAYERS_FP32, LAYERS_TENSOR, LSTM_FAITHFUL, LAYERS_FAITHFUL)

__all__ = ["CHUNK", "SAMPLE_RATE", "Engine", "Group", "EngineError", "StreamSegmenter", "lib", "pinned_empty", "pinned_free",
           "seg_params", "segments_text", "synth_pcm", "LIB_PATH", "WEIGHTS_PATH", "STFT_AUTO", "STFT_HYBRID", "STFT_EXACT",
           "LSTM_AUTO", "LSTM_FP32", "LSTM_TENSOR", "LAYERS_AUTO", "LAYERS_FP32", "LAYERS_TENSOR", "LSTM_FAITHFUL", "LAYERS_FAITHFUL"]
