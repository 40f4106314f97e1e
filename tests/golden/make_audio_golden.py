"""Generates tests/golden/audio_golden.npz by running the REFERENCE's own AudioSegment.normalize / to('int16')
(/root/reference/ppasr/data_utils/audio.py:244-304,519-574 -- NumPy only). audio.py imports resampy / soundfile and
ppasr.data_utils.utils at module level for code paths that are not exercised here (file decoding, resampling); those modules
are absent in this sandbox, so they are replaced by empty stub modules before the file is loaded. The arithmetic that runs is the
reference's own.

Run once in the build container (where /root/reference exists):  python tests/golden/make_audio_golden.py
The committed .npz is what the CPU tests compare oracle/fbank_oracle.py (db_normalize, to_int16_scale) against.
"""
import importlib.util
import os
import sys
import types

import numpy as np

if not hasattr(np, "sctypes"):  # removed in NumPy 2; audio.py:542,567 use np.sctypes['float']
    np.sctypes = {"float": [np.float16, np.float32, np.float64], "int": [np.int8, np.int16, np.int32, np.int64]}
for name in ("resampy", "soundfile", "ppasr", "ppasr.data_utils", "ppasr.data_utils.utils"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["ppasr.data_utils.utils"].buf_to_float = None
sys.modules["ppasr.data_utils.utils"].decode_audio = None

REF = "/root/reference/ppasr/data_utils/audio.py"
spec = importlib.util.spec_from_file_location("ref_audio", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.RandomState(7)
out = {}
cases = {
    "quiet": (rng.randn(4000) * 0.003).astype(np.float32),
    "loud_clipping": np.clip(rng.randn(4000) * 0.8, -1, 1).astype(np.float32),
    "tone": (0.25 * np.sin(2 * np.pi * 440 * np.arange(8000) / 16000.0)).astype(np.float32),
    "int16_input": (rng.randn(3000) * 2000).astype(np.int16),
    "silence": np.zeros(1000, dtype=np.float32),
}
for k, x in cases.items():
    seg = ref.AudioSegment(x, 16000)
    out[k + "/input"] = x
    out[k + "/float32"] = seg.samples.copy() if hasattr(seg, "samples") else seg._samples.copy()
    out[k + "/rms_db"] = np.float64(seg.rms_db)
    seg.normalize(target_db=-20)
    out[k + "/normalized"] = seg._samples.copy()
    out[k + "/int16"] = seg.to("int16")
np.savez_compressed(os.path.join(HERE, "audio_golden.npz"), **out)
print("wrote", len(out), "arrays")
