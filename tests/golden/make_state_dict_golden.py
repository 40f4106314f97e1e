"""Generate tests/golden/state_dict_golden.json: the parameter names and shapes of the reference's OWN model classes built from
the reference's OWN shipped configs (configs/{conformer,squeezeformer,efficient_conformer,deepspeech2}.yml), the way
PPASRTrainer builds them (trainer.py:172-208): <Family>Model(input_dim, vocab_size, mean_istd_path, streaming, encoder_conf,
decoder_conf, **model_conf). Both settings of `streaming` are recorded (it selects causal convolutions, the Squeezeformer
time-reduction layer and the DeepSpeech2 RNN direction). `paddle` is tests/golden/paddle_shim; training-only imports (av,
zhconv, resampy, soundfile, termcolor, paddleaudio, paddle.io) are empty stand-ins.

The CPU test test_host_cpu.py::test_param_tables_match_reference_state_dict checks that ppasr_b200/weights.py expects exactly
these names and shapes for the inference path (everything but the attention decoder `decoder.*` of the *former models, which
CTC inference never reads). Run in the build container only:  python tests/golden/make_state_dict_golden.py
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import yaml

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "paddle_shim"))
sys.path.insert(0, "/root/reference")

if not hasattr(np, "sctypes"):
    np.sctypes = {"float": [np.float16, np.float32, np.float64], "int": [np.int8, np.int16, np.int32, np.int64]}
for name in ("av", "zhconv", "resampy", "soundfile", "termcolor", "paddleaudio", "paddleaudio.compliance",
             "paddleaudio.compliance.kaldi", "paddle.io"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["zhconv"].convert = lambda s, *_: s
sys.modules["termcolor"].colored = lambda s, *a, **k: s
sys.modules["paddleaudio.compliance.kaldi"].fbank = None
sys.modules["paddleaudio.compliance.kaldi"].mfcc = None
sys.modules["paddle.io"].Dataset = object
sys.modules["paddle.io"].DataLoader = object

import paddle  # noqa: E402,F401  (shim)

from ppasr.model_utils.conformer.model import ConformerModel  # noqa: E402
from ppasr.model_utils.deepspeech2.model import DeepSpeech2Model  # noqa: E402
from ppasr.model_utils.efficient_conformer.model import EfficientConformerModel  # noqa: E402
from ppasr.model_utils.squeezeformer.model import SqueezeformerModel  # noqa: E402

CLASSES = {"conformer": ConformerModel, "squeezeformer": SqueezeformerModel, "efficient_conformer": EfficientConformerModel,
           "deepspeech2": DeepSpeech2Model}
VOCAB, N_MELS = 4233, 80


def main():
    d = tempfile.mkdtemp()
    mi = os.path.join(d, "mean_istd.json")
    json.dump({"mean": [0.0] * N_MELS, "istd": [1.0] * N_MELS}, open(mi, "w"))
    out = {"vocab_size": VOCAB, "n_mels": N_MELS, "models": []}
    for name, cls in CLASSES.items():
        cfg = yaml.safe_load(open(f"/root/reference/configs/{name}.yml", encoding="utf-8"))
        assert cfg["use_model"] == name
        for streaming in (True, False):
            kw = {} if name == "deepspeech2" else dict(cfg.get("model_conf") or {})
            m = cls(input_dim=N_MELS, vocab_size=VOCAB, mean_istd_path=mi, streaming=streaming,
                    encoder_conf=cfg["encoder_conf"], decoder_conf=cfg["decoder_conf"], **kw)
            sd = {k: [int(v) for v in t.shape] for k, t in m.state_dict().items()}
            att_dec = {k: v for k, v in sd.items() if name != "deepspeech2" and k.startswith("decoder.")}
            out["models"].append({
                "use_model": name, "streaming": streaming, "yml_streaming": bool(cfg["streaming"]),
                "encoder_conf": cfg["encoder_conf"],
                "state_dict": {k: v for k, v in sd.items() if k not in att_dec},
                "attention_decoder_tensors": len(att_dec),
            })
            print(name, streaming, len(sd), "tensors,", len(att_dec), "in the attention decoder")
    with open(os.path.join(HERE, "state_dict_golden.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)


if __name__ == "__main__":
    main()
