#!/usr/bin/env python
"""Generate tests/golden/cli/: FASTA / matrix fixtures (own, seeded synthetic data) and the stdout the UNMODIFIED
reference CLI (oracle/_ref/opal_aligner_ref, built from /root/reference/src by oracle/Makefile) prints for them.

Run:  make -C oracle && python tests/golden/make_cli_golden.py
The GPU box has no /root/reference; tests/test_gpu_cli.py only reads the committed files and compares the output of
opal_aligner_b200 with them line by line (the timing lines excepted).
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from opal_b200 import datasets, matrices  # noqa: E402

OUT = os.path.join(HERE, "cli")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "opal_aligner_ref")

# option sets run on (query.fasta, db) pairs; "messy" fixtures need an alphabet with '*', i.e. the built-in matrix
CASES = {
    "sw_score": (["-s"], "query.fasta", "db_messy.fasta"),
    "sw_default": ([], "query.fasta", "db_messy.fasta"),
    "sw_end": (["-x", "1"], "query.fasta", "db_messy.fasta"),
    "sw_align": (["-x", "2"], "query.fasta", "db_messy.fasta"),
    "sw_align_gaps": (["-x", "2", "-o", "5", "-e", "2"], "query.fasta", "db_messy.fasta"),
    "nw_end_b62": (["-a", "NW", "-x", "1", "-o", "11", "-e", "1", "-f", "blosum62.mat"], "query.fasta", "db_clean.fasta"),
    "hw_end_b62": (["-a", "HW", "-x", "1", "-o", "11", "-e", "1", "-f", "blosum62.mat"], "query.fasta", "db_clean.fasta"),
    "ov_end_b62": (["-a", "OV", "-x", "1", "-o", "11", "-e", "1", "-f", "blosum62.mat"], "query.fasta", "db_clean.fasta"),
    "nw_align_b62": (["-a", "NW", "-x", "2", "-o", "11", "-e", "1", "-f", "blosum62.mat"], "query.fasta", "db_clean.fasta"),
    "hw_align_b62": (["-a", "HW", "-x", "2", "-o", "11", "-e", "1", "-f", "blosum62.mat"], "query.fasta", "db_clean.fasta"),
    "ov_align_b62": (["-a", "OV", "-x", "2", "-o", "11", "-e", "1", "-f", "blosum62.mat"], "query.fasta", "db_clean.fasta"),
    "two_queries_first_only": (["-x", "1"], "queries2.fasta", "db_clean.fasta"),
}


def letters(codes, sm):
    return "".join(sm.alphabet[int(c)] for c in codes)


def wrap(text, width):
    return "\n".join(text[i:i + width] for i in range(0, len(text), width))


def write_fixtures():
    rng = np.random.default_rng(20261020)
    sm = matrices.blosum62()
    q = datasets.random_residues(93, rng, sm)
    q2 = datasets.random_residues(41, rng, sm)
    seqs = []
    for k in range(26):
        n = int(rng.integers(25, 420))
        s = datasets.random_residues(n, rng, sm)
        if k % 4 == 1:
            s = datasets.mutate(q, float(rng.uniform(0.45, 0.95)), rng, sm)
        if k % 9 == 5:  # a homolog embedded in unrelated flanks
            s = np.concatenate([datasets.random_residues(60, rng, sm), datasets.mutate(q, 0.8, rng, sm), datasets.random_residues(35, rng, sm)])
        seqs.append(s)
    with open(os.path.join(OUT, "query.fasta"), "w") as f:
        f.write(">sp|Q00001|QUERY_SYNTH synthetic query\n" + wrap(letters(q, sm), 60) + "\n")
    with open(os.path.join(OUT, "queries2.fasta"), "w") as f:
        f.write(">first\n" + wrap(letters(q, sm), 70) + "\n>second\n" + wrap(letters(q2, sm), 70) + "\n")
    with open(os.path.join(OUT, "db_clean.fasta"), "w") as f:
        for k, s in enumerate(seqs):
            f.write(f">sp|T{k:05d}|TARGET_{k} synthetic target {k}\n" + wrap(letters(s, sm), 60) + "\n")
    # the same database written carelessly: CRLF line ends, ragged line lengths, lower-case and non-standard
    # letters (all -> '*' under the built-in alphabet), headers without a sequence, no newline at the end
    with open(os.path.join(OUT, "db_messy.fasta"), "wb") as f:
        f.write(b">empty record at the start\r\n>another > one\n")
        for k, s in enumerate(seqs):
            text = letters(s, sm)
            if k % 5 == 2:
                text = text[:7] + text[7:19].lower() + text[19:]
            if k % 7 == 3:
                text = text[:11] + "UOJ" + text[11:]
            eol = b"\r\n" if k % 3 == 0 else b"\n"
            width = [60, 80, 13, 200][k % 4]
            f.write(f">T{k} messy".encode() + eol)
            f.write(eol.join(wrap(text, width).encode().split(b"\n")))
            f.write(eol if k + 1 < len(seqs) else b"")
            if k == 10:
                f.write(b">no residues here\n\n")
    with open(os.path.join(OUT, "blosum62.mat"), "w") as f:
        f.write(" ".join(sm.alphabet) + "\n")
        for row in sm.matrix:
            f.write(" ".join(str(int(v)) for v in row) + "\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    write_fixtures()
    index = {}
    for name, (opts, qf, dbf) in CASES.items():
        opts2 = [os.path.join(OUT, o) if o.endswith(".mat") else o for o in opts]
        p = subprocess.run([REF_CLI] + opts2 + [os.path.join(OUT, qf), os.path.join(OUT, dbf)], capture_output=True, text=True)
        assert p.returncode == 0, (name, p.returncode, p.stderr[-400:])
        with open(os.path.join(OUT, name + ".txt"), "w") as f:
            f.write(p.stdout)
        index[name] = {"options": opts, "query": qf, "db": dbf}
        print(name, len(p.stdout), "bytes")
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
