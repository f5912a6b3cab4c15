import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import vadc_b200
e = vadc_b200.Engine()
x = np.maximum(np.random.default_rng(1).standard_normal((3000, 7, 64)).astype(np.float32), 0)
wave = len(sys.argv) > 1 and sys.argv[1] == "wave"
e.stage_exact_lstm(x[:50], wave=wave)
e.stage_exact_lstm(x, wave=wave)
