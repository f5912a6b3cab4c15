"""GPU: the native Linux CLI (vadc_b200/vadc_b200_cli) against the unmodified
reference CLI built for Linux (oracle/_ref/vadc_linux): same stdin contract, same options, same stdout bytes."""
import os
import subprocess

import numpy as np
import pytest

import vadc_b200
from oracle_lib import ROOT, have_ref, ref_cli

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "vadc_b200", "vadc_b200_cli")


def cli(pcm, *args, files=()):
    data = np.ascontiguousarray(pcm, np.int16).tobytes() if pcm is not None else b""
    r = subprocess.run([CLI, *args, *files], input=data, capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    return r.stdout.decode()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("args", [(), ("--output_centi_seconds",), ("--batch", "32"), ("--batch", "1"),
                                  ("--threshold", "0.6", "--min_silence", "300", "--speech_pad", "100"),
                                  ("--min_speech", "-5", "--neg_threshold_relative", "0.3")])
def test_stdin_mode_reproduces_reference_cli_stdout(args):
    pcm = vadc_b200.synth_pcm(515, 16000 * 45 + 700)      # 45 s + a partial trailing chunk (dropped, vadc.c:964)
    want = ref_cli(pcm, *args)
    assert want.count("\n") >= 5
    assert cli(pcm, *args) == want


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_raw_probabilities_and_edge_inputs():
    pcm = vadc_b200.synth_pcm(99, 16000 * 20)
    got = np.array([float(v) for v in cli(pcm, "--raw_probabilities").split()])
    ref = np.array([float(v) for v in ref_cli(pcm, "--raw_probabilities").split()])
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4 + 1e-6
    assert cli(np.zeros(0, np.int16)) == ref_cli(np.zeros(0, np.int16)) == ""
    assert cli(np.zeros(1000, np.int16)) == ""               # less than one chunk
    loud = vadc_b200.synth_pcm(5, 16000 * 10, kind=2)        # full-scale white noise
    assert cli(loud) == ref_cli(loud)


def test_files_are_concurrent_streams(tmp_path):
    """Each file = one stream on the multi-stream scheduler + device segmenter; ragged lengths, one empty file, one
    whose length is an exact multiple of the call size (closed by the end-of-stream flush)."""
    lengths = [64, 50, 7, 0, 32, 50]                        # chunks; --batch 2 -> 32 chunks per call
    paths, pcms = [], []
    for i, n in enumerate(lengths):
        pcm = vadc_b200.synth_pcm(900 + i, n * 1536 + (300 if i == 1 else 0))
        p = tmp_path / ("s%d.s16le" % i)
        p.write_bytes(pcm.tobytes())
        paths.append(str(p))
        pcms.append(pcm)
    out = cli(None, "--batch", "2", files=paths)
    want = "".join("# %s\n%s" % (p, cli(pcm)) for p, pcm in zip(paths, pcms))
    assert out == want
    assert want.count(",") >= 4
    one = cli(None, files=paths[:1])                        # a single file prints no header
    assert one == cli(pcms[0])
    raw = cli(None, "--raw_probabilities", "--batch", "2", files=paths[1:3])
    vals = [l for l in raw.splitlines() if not l.startswith("#")]
    assert len(vals) == 50 + 7


def _fake_ffmpeg(tmp_path):
    """A stand-in for ffmpeg (absent from this image): logs its argument vector, honours -ss on a raw s16le 'container'
    and writes mono 16 kHz s16le to stdout -- what the real decoder's output contract is (vadc.c:537)."""
    script = tmp_path / "ffmpeg"
    script.write_text(
        "#!/bin/bash\n"
        "printf '%s\\n' \"$@\" > \"$FAKE_FFMPEG_LOG.$$\"\n"
        "ss=0; inp=\n"
        "while [ $# -gt 0 ]; do case \"$1\" in -ss) ss=$2; shift;; -i) inp=$2; shift;; esac; shift; done\n"
        "skip=$(python3 -c \"print(int(float('$ss')*16000)*2)\")\n"
        "tail -c +$((skip+1)) \"$inp\"\n")
    script.chmod(0o755)
    return str(script)


def test_named_input_goes_through_an_ffmpeg_child(tmp_path):
    """vadc.c:531-626: a named input is decoded by `ffmpeg ... -ss S -i FILE -map 0:a:N ... -f s16le -`; same stdout as
    piping the decoded samples, the reference's argument vector, --start_seconds as -ss, several inputs = several streams."""
    pcm = vadc_b200.synth_pcm(321, 16000 * 30 + 99)
    media = tmp_path / "talk.wav"              # any name that is not *.s16le/*.raw/*.pcm takes the decoder path
    media.write_bytes(pcm.tobytes())
    env = dict(os.environ, VADC_FFMPEG=_fake_ffmpeg(tmp_path), FAKE_FFMPEG_LOG=str(tmp_path / "argv"))

    def run(*args):
        r = subprocess.run([CLI, *args], stdin=subprocess.DEVNULL, capture_output=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr.decode()
        return r.stdout.decode()

    assert run(str(media)) == cli(pcm)
    logs = sorted(p for p in os.listdir(tmp_path) if p.startswith("argv."))
    argv = (tmp_path / logs[-1]).read_text().split("\n")[:-1]
    assert argv == ["-hide_banner", "-loglevel", "error", "-nostats", "-ss", "0.000000", "-i", str(media), "-map", "0:a:0",
                    "-vn", "-sn", "-dn", "-ac", "1", "-ar", "16k", "-f", "s16le", "-"]
    assert run("--start_seconds", "7.5", "--audio_source", "2", str(media)) == cli(pcm[int(7.5 * 16000):])
    newest = max((tmp_path / p for p in os.listdir(tmp_path) if p.startswith("argv.")), key=lambda p: p.stat().st_mtime)
    argv = newest.read_text().split("\n")
    assert argv[4:6] == ["-ss", "7.500000"] and argv[8:10] == ["-map", "0:a:2"]
    # decoded and raw inputs mixed: every file is a stream of the multi-stream scheduler
    raw = tmp_path / "other.s16le"
    pcm2 = vadc_b200.synth_pcm(322, 16000 * 12)
    raw.write_bytes(pcm2.tobytes())
    assert run(str(media), str(raw)) == "# %s\n%s# %s\n%s" % (media, cli(pcm), raw, cli(pcm2))
    # a decoder that cannot be launched: diagnostics on stderr, no segments
    r = subprocess.run([CLI, str(media)], stdin=subprocess.DEVNULL, capture_output=True, timeout=300, env=dict(env, VADC_FFMPEG="/nonexistent/ffmpeg"))
    assert r.stdout == b"" and b"Error launching ffmpeg" in r.stderr


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_stats_line_matches_the_reference_cli():
    """--stats (vadc.c:1037-1081): "time=HH:MM:SS.mmmm  <speech> speech (<pct>%), <duration> / <elapsed> (<speed>x)" on stderr, rewritten
    with \\r after every batch and ended with a newline. Speech total, percentage and duration are functions of the segments and must
    equal the reference's; elapsed time and speed are the run's own. Stdout stays the segment list. And with --start_seconds on stdin
    (which the reference hands to ffmpeg only) the output does not change."""
    import re
    from oracle_lib import REF_CLI
    pcm = vadc_b200.synth_pcm(321, 16000 * 60)
    data = pcm.tobytes()
    pat = re.compile(r"time=(\d\d:\d\d:\d\d\.\d{4})\s+([\d.]+) speech \(\s*([\d.]+)%\),\s+([\d.]+) /\s+([\d.]+) \(\s*([\d.]+)x\)")
    outs = {}
    for name, exe in (("ours", CLI), ("ref", REF_CLI)):
        r = subprocess.run([exe, "--stats"], input=data, capture_output=True, timeout=300)
        assert r.returncode == 0
        lines = [m for m in pat.finditer(r.stderr.decode(errors="replace"))]
        assert lines, r.stderr[-300:]
        outs[name] = (r.stdout, lines[-1].groups(), len(lines))
    assert outs["ours"][0] == outs["ref"][0]
    assert outs["ours"][1][1:4] == outs["ref"][1][1:4]      # speech seconds, percentage, duration
    # the time field: the reference counts what its buffered reader reports per refill (vadc.c:861-864), which on the Linux shim of the
    # Win32 reader runs a few milliseconds past the audio at end of file; this program counts the chunks it processed
    t = lambda v: int(v[0:2]) * 3600 + int(v[3:5]) * 60 + int(v[6:8]) + int(v[9:13]) / 1000.0
    assert abs(t(outs["ours"][1][0]) - t(outs["ref"][1][0])) <= 0.1 and t(outs["ours"][1][0]) == 60.0
    assert outs["ours"][2] >= 2                              # progress lines (\r) + the final one (the reference also reprints per segment)
    shifted = subprocess.run([CLI, "--start_seconds", "7"], input=data, capture_output=True, timeout=300)
    assert shifted.stdout == outs["ours"][0]


def test_multi_file_stats_count_speech(tmp_path):
    """Multi-file mode: the final --stats line sums the speech of all files (round 1 printed 0)."""
    import re
    paths = []
    for i in range(3):
        p = tmp_path / ("s%d.s16le" % i)
        vadc_b200.synth_pcm(70 + i, 200 * 1536).tofile(p)
        paths.append(str(p))
    r = subprocess.run([CLI, "--stats"] + paths, capture_output=True, timeout=300)
    assert r.returncode == 0
    m = re.findall(r"([\d.]+) speech \(", r.stderr.decode(errors="replace"))
    assert m and float(m[-1]) > 1.0


def test_pipe_in_pieces_and_file_input_give_the_same_output(tmp_path):
    """The reader thread (next batch read while the GPU works) and the larger calls on regular-file input change when bytes are
    read, not what is computed: a pipe fed in small, slow pieces, a regular file as stdin (calls of 1536 chunks), the same file with
    an explicit --batch, and the one-shot pipe of the other tests all print the same segments."""
    import time
    pcm = vadc_b200.synth_pcm(321, 1536 * 3300 + 555)     # more than two 1536-chunk calls + a partial trailing chunk
    data = pcm.tobytes()
    want = cli(pcm)
    assert want.count("\n") >= 20
    path = tmp_path / "in.s16le"
    path.write_bytes(data)
    for args in ((), ("--batch", "96"), ("--batch", "1000")):
        with open(path, "rb") as f:
            r = subprocess.run([CLI, *args], stdin=f, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout.decode() == want, args
    p = subprocess.Popen([CLI, "--batch", "7"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    step = 1536 * 2 * 5 + 123                             # pieces that straddle batch boundaries
    for i in range(0, min(len(data), step * 40), step):
        p.stdin.write(data[i:i + step])
        p.stdin.flush()
        time.sleep(0.002)
    p.stdin.write(data[step * 40:])
    out, err = p.communicate(timeout=300)
    assert p.returncode == 0, err.decode()
    assert out.decode() == want
