"""Instruction-length histogram of the reference's shipped R2R training set (data/R2R_train.json, tokenised with the
reference's vocabulary and rules, encoding length 80): the distribution `make_items(length_hist=...)` / `bench.py
--lengths real` draw synthetic instruction lengths from (SURVEY 8d "a second run draws lengths from the real R2R
distribution").  TEST / BENCH INFRASTRUCTURE: reads /root/reference, writes a 81-entry count vector; run once here.
usage: python oracle/make_length_hist.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402,F401
from clvln_b200.environ import ingest  # noqa: E402
from oracle import ref_loader  # noqa: E402

data_dir = os.path.join(ref_loader.REF_TASK, "data")
tok = ingest.Tokenizer.from_file(os.path.join(data_dir, "train_vocab.txt"), 80)     # pinned == reference Tokenizer
hist = [0] * 81
with open(os.path.join(data_dir, "R2R_train.json")) as f:
    for item in json.load(f):
        for instr in item["instructions"]:
            enc = tok.encode_sentence(instr)
            if enc is not None:
                hist[enc[1]] += 1
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curriculum-learning-for-vln_b200", "environ",
                   "r2r_train_lengths.json")
n = sum(hist)
mean = sum(i * c for i, c in enumerate(hist)) / n
json.dump({"source": "data/R2R_train.json, Tokenizer(train_vocab.txt, encoding_length=80): count of instructions per instr_length",
           "n": n, "mean": round(mean, 2), "counts": hist}, open(out, "w"))
print(n, round(mean, 2), "max", max(i for i, c in enumerate(hist) if c), "at80", hist[80])
