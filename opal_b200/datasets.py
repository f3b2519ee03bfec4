"""Synthetic workloads for tests and bench.py (caller-side data, not the hot path).

The reference's benchmark databases are not shipped (reference
.MISSING_LARGE_BLOBS:1; test/perf:15-24 point into the author's home), so the
BASELINE.json configs are regenerated here with one seeded numpy generator, as
SURVEY.md section 8d specifies: Swiss-Prot-like log-normal lengths, Swiss-Prot
background residue frequencies, planted homologs of the query so that scores
span the 16-bit boundary, and a heavy tail where a config asks for one.
Sequences are arrays of alphabet indices (reference src/opal.h:96-98).
"""
from __future__ import annotations

import numpy as np

from .capi import SequenceDB
from .matrices import ScoreMatrix

# UniProt P18080 (513 aa) and O74807 (110 aa): the two query proteins BASELINE.json names
# (reference test_data/query/P18080.fasta, O74807.fasta are the same public records).
P18080 = (
    "MAAFLRCPLLARHPPLARAFATGARCPFMGFAHRAAPELQEDVERPQIPAVEVLEELLRDGGAALNRTVRDCMDEDAFPYEEQFQAQLGALRRTHTYRVV"
    "TAVGRRADAPPLGTRGTAPHTSVELWCSSDYLGLSRHPAVLRAARAALDAHGLGAGGTRNIGGTSPLHGALERALALLHRQPRAALFSSCFAANDTALDT"
    "LARILPGCQVYSDAGNHASMIQGIRRRGVPKFIFRHNDPHHLEQLLGRSPPGVPKIVAFESLHSMDGSIAPLEELCDVAHAYGALTFVDEVHAVGLYGAR"
    "GAGIAERDGVQHKVDVVSGTLGKALGAVGGYIAGSEALVDAVRSLGPGFIFTTALPPQRGGGALAALQVVGSAEGAALRRAHQRHAKHLRVLLRDRGLPA"
    "LPSHIVPVRWDAEANTRLSRALLEEHGLYVQAINHPTVPRGQELLLRIAPTPHHSPPMLENLADKLSECWGAVGLPREDPPGPSCSSCHRPLHLSLLSPL"
    "ERDQFGVRGAAAG"
)
O74807 = (
    "MMEEERFKAEIFHVTQEVCNRTASELTESESRNVIVDELFCVGVTEMVWEQIRVLAKDIEAFAEHAGRKTVQPQDVLLCCRRNEGLYEIINNFHKESIKS"
    "KKKKKENSTT"
)

# Swiss-Prot background composition in percent.
_AA_PERCENT = {
    "L": 9.65, "A": 8.25, "G": 7.07, "V": 6.86, "E": 6.72, "S": 6.65, "I": 5.91, "K": 5.80, "R": 5.53,
    "D": 5.46, "T": 5.36, "P": 4.74, "N": 4.06, "Q": 3.93, "F": 3.86, "Y": 2.92, "M": 2.41, "H": 2.27,
    "C": 1.38, "W": 1.10,
}
# The customary 20 query lengths of the SWIPE / CUDASW++ protocol.
CONFIG3_QUERY_LENGTHS = [144, 189, 222, 375, 464, 567, 657, 729, 850, 1000, 1500, 2005, 2504, 3005,
                         3564, 4061, 4548, 4743, 5147, 5478]

LOGN_MU, LOGN_SIGMA = 5.687, 0.635
MAX_PROTEIN_LEN = 35213


def residue_distribution(sm: ScoreMatrix):
    """(codes, probabilities) of the 20 standard amino acids in `sm`'s alphabet."""
    codes = np.array([sm.alphabet.index(a) for a in _AA_PERCENT], dtype=np.uint8)
    p = np.array(list(_AA_PERCENT.values()), dtype=np.float64)
    return codes, p / p.sum()


def random_residues(n, rng, sm: ScoreMatrix):
    codes, p = residue_distribution(sm)
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    return codes[np.searchsorted(cdf, rng.random(n), side="right")]


def lognormal_lengths(n, rng, lo=2, hi=MAX_PROTEIN_LEN):
    return np.clip(np.rint(rng.lognormal(LOGN_MU, LOGN_SIGMA, n)), lo, hi).astype(np.int64)


def mutate(seq, identity, rng, sm: ScoreMatrix, indel_rate=0.03):
    """A homolog of `seq`: each residue kept with probability `identity`, else substituted;
    short insertions/deletions at `indel_rate` per position."""
    out = []
    i = 0
    n = len(seq)
    while i < n:
        u = rng.random()
        if u < indel_rate / 2:            # deletion of 1-5 residues
            i += int(rng.integers(1, 6))
            continue
        if u < indel_rate:                # insertion of 1-5 residues
            out.extend(random_residues(int(rng.integers(1, 6)), rng, sm).tolist())
        out.append(int(seq[i]) if rng.random() < identity else int(random_residues(1, rng, sm)[0]))
        i += 1
    return np.array(out if out else [int(seq[0])], dtype=np.uint8)


def _assemble(lengths, rng, sm, planted=None):
    """Concatenated residue buffer for `lengths`; `planted` maps index -> explicit sequence."""
    lengths = np.asarray(lengths, dtype=np.int64).copy()
    planted = planted or {}
    for i, s in planted.items():
        lengths[i] = len(s)
    offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    residues = random_residues(int(offsets[-1]), rng, sm)
    for i, s in planted.items():
        residues[offsets[i]:offsets[i + 1]] = s
    return SequenceDB(residues, offsets)


def protein_db(n, seed, sm: ScoreMatrix, query=None, homolog_fraction=0.01, tail_fraction=0.0,
               exact_max=False, residue_seed=None):
    """Swiss-Prot-shaped synthetic protein database.

    n sequences with log-normal lengths; `homolog_fraction` of them are mutated copies of
    `query` (30-90 % identity, with indels); `tail_fraction` of them are redrawn log-uniform in
    [5000, 35213]; with `exact_max` one sequence has exactly 35213 residues.  `residue_seed` redraws the
    background residues only: same lengths and homologs, different sequences (equal-work shards for weak scaling).
    """
    rng = np.random.default_rng(seed)
    lengths = lognormal_lengths(n, rng)
    if tail_fraction > 0:
        k = max(1, int(round(n * tail_fraction)))
        idx = rng.choice(n, size=k, replace=False)
        lengths[idx] = np.exp(rng.uniform(np.log(5000), np.log(MAX_PROTEIN_LEN), k)).astype(np.int64)
        if exact_max:
            lengths[idx[0]] = MAX_PROTEIN_LEN
    planted = {}
    if query is not None and homolog_fraction > 0:
        k = max(1, int(round(n * homolog_fraction)))
        for i in rng.choice(n, size=k, replace=False):
            ident = rng.uniform(0.3, 0.9)
            core = mutate(query, ident, rng, sm)
            left = random_residues(int(rng.integers(0, 60)), rng, sm)
            right = random_residues(int(rng.integers(0, 60)), rng, sm)
            planted[int(i)] = np.concatenate([left, core, right]).astype(np.uint8)
    return _assemble(lengths, rng if residue_seed is None else np.random.default_rng(residue_seed), sm, planted)


def config2_db(sm: ScoreMatrix, query, n=12071, seed=20261017, residue_seed=None):
    """BASELINE.json configs[1]: 12,071-sequence Swiss-Prot-length-distributed DB (~4.3 M residues)."""
    return protein_db(n, seed, sm, query=query, homolog_fraction=0.01, residue_seed=residue_seed)


def config3_db(sm: ScoreMatrix, n=570000, seed=20261018, query=None):
    """BASELINE.json configs[2]: 570k sequences / ~206 M residues with a planted heavy tail."""
    return protein_db(n, seed, sm, query=query, homolog_fraction=0.0 if query is None else 0.001,
                      tail_fraction=0.0002, exact_max=True)


def config3_queries(sm: ScoreMatrix, seed=20261018):
    rng = np.random.default_rng(seed + 1)
    return [random_residues(L, rng, sm) for L in CONFIG3_QUERY_LENGTHS]


def dna_db(n, seed, query=None, alpha=1.2, xmin=200, max_len=100000, n_at_max=10, n_planted=4):
    """BASELINE.json configs[4]: DNA (alphabet 4), Pareto(alpha, xmin) lengths clipped to max_len,
    at least `n_at_max` sequences of exactly max_len, a few long high-identity copies of the query."""
    rng = np.random.default_rng(seed)
    lengths = np.minimum((xmin * (1.0 - rng.random(n)) ** (-1.0 / alpha)).astype(np.int64), max_len)
    if n_at_max and n >= n_at_max:
        lengths[rng.choice(n, size=n_at_max, replace=False)] = max_len
    offsets = np.zeros(n + 1, dtype=np.int64)
    planted = {}
    if query is not None and n_planted:
        for i in rng.choice(n, size=min(n_planted, n), replace=False):
            keep = rng.random(len(query)) < 0.97
            s = np.where(keep, query, rng.integers(0, 4, len(query))).astype(np.uint8)
            planted[int(i)] = s
            lengths[int(i)] = len(s)
    np.cumsum(lengths, out=offsets[1:])
    residues = rng.integers(0, 4, int(offsets[-1]), dtype=np.uint8)
    for i, s in planted.items():
        residues[offsets[i]:offsets[i + 1]] = s
    return SequenceDB(residues, offsets)


def read_fasta(path, sm: ScoreMatrix):
    """FASTA -> list of index arrays (the job of reference src/opal_aligner.cpp:247-301)."""
    seqs, cur = [], None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if cur is not None:
                    seqs.append(sm.encode("".join(cur)))
                cur = []
            elif cur is not None:
                cur.append(line)
    if cur is not None:
        seqs.append(sm.encode("".join(cur)))
    return seqs
