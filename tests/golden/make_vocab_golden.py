"""Generates tests/golden/vocab_golden.json by running the REFERENCE's own TextFeaturizer
(/root/reference/ppasr/data_utils/featurizer/text_featurizer.py -- plain Python, no imports) on a small vocabulary file.
Run once in the build container:  python tests/golden/make_vocab_golden.py"""
import importlib.util
import json
import os
import tempfile

REF = "/root/reference/ppasr/data_utils/featurizer/text_featurizer.py"
spec = importlib.util.spec_from_file_location("ref_text", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

HERE = os.path.dirname(os.path.abspath(__file__))
content = "<blank>\t-1\n<unk>\t-1\n的\t1203\n一\t877\n <space>\t12\nab\t3\n是\t650\n<eos>\t-1\n"
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "vocabulary.txt")
    with open(p, "w", encoding="utf-8") as f:
        f.write(content)
    tf = ref.TextFeaturizer(p)
    out = {"file": content, "vocab_list": tf.vocab_list, "vocab_size": tf.vocab_size,
           "featurize": {t: tf.featurize(t) for t in ["的一是", "一 是", "x的"]}}
json.dump(out, open(os.path.join(HERE, "vocab_golden.json"), "w", encoding="utf-8"), ensure_ascii=False, indent=1)
print(out["vocab_list"], out["featurize"])
