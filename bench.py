#!/usr/bin/env python
"""bench.py -- GCUPS of the database-search hot path on N B200s, beside the reference on host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload config2|config3]

A step is one pass of the hot path: one query against one resident database (per GPU).  The
default workload is BASELINE.json configs[1]: SW score+end, query P18080 (513 aa) against a
synthetic 12,071-sequence Swiss-Prot-length-distributed database, BLOSUM62, gaps 11/1.  With
N > 1 (one process per GPU under torchrun) every rank holds its own 12,071-sequence shard of an
N x 12,071-sequence database (weak scaling, no data-path collective: targets are independent).
`--workload config3` runs the 570k-sequence / 206 M-residue database of configs[2] instead,
dealt residue-balanced over the ranks (strong scaling).

Metric: GCUPS = queryLength * sum(dbSeqLengths) / 1e9 / seconds (reference
src/opal_aligner.cpp:205-206).  `value` is device-timed (CUDA events on the library's stream, first
kernel launch to last kernel end, database resident, max over ranks); `e2e` goes through the
reference-facing C ABI call opalSearchDatabase with host buffers, so database packing, H2D, D2H
and the per-record result writes are all inside its timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from opal_b200 import (MODES, OPAL_OVERFLOW_BUCKETS, OPAL_SEARCH_SCORE_END, OpalCLibrary, SequenceDB,  # noqa: E402
                       datasets, matrices, new_results, sharding)

GAP_OPEN, GAP_EXT = 11, 1
MODE = "SW"


# ----------------------------------------------------------------------------- workloads
def make_workload(name, rank, world):
    sm = matrices.blosum62()
    query = sm.encode(datasets.P18080)
    if name == "config2":
        # weak scaling: every GPU gets a shard of exactly the same work -- the same lengths and planted homologs, the
        # background residues redrawn per rank -- so that the per-GPU tail (the longest target) does not vary by rank
        db = datasets.config2_db(sm, query, residue_seed=None if rank == 0 else 20261017 + rank)
        desc = ("BASELINE configs[1]: SW score+end, P18080 (Q=513) vs synthetic 12,071-seq Swiss-Prot-shaped DB, "
                "BLOSUM62 11/1; one such shard per GPU (same lengths, residues redrawn per rank)")
        scaling = "weak"
    elif name == "config3":
        full = datasets.config3_db(sm, query=query)
        db = sharding.shard_db(full, sharding.deal_shards(full.lengths, world)[rank]) if world > 1 else full
        desc = ("BASELINE configs[2] DB: 570k seqs / ~206M residues, Swiss-Prot-shaped with heavy tail, "
                "SW score+end, P18080 (Q=513), BLOSUM62 11/1; DB dealt residue-balanced over the GPUs")
        scaling = "strong"
    else:
        raise SystemExit(f"unknown workload {name}")
    return sm, query, db, desc, scaling


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.ok = [], set(), False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.005)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        self.thread.join()
        self.sample()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU reference arm
def ncu_dram_bytes_per_launch():
    """DRAM bytes of the dominant kernel launch from the committed ncu capture (None if it is not there)."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_search_kernel.txt")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    try:
        cur = {}
        for line in open(path):
            tok = line.split()
            if len(tok) >= 3 and tok[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and tok[1] in unit:
                cur[tok[0]] = float(tok[2].replace(",", "")) * unit[tok[1]]
                if len(cur) == 2:
                    total = sum(cur.values())
                    best = total if best is None else max(best, total)
                    cur = {}
        return best
    except OSError:
        return None


def cpu_library():
    ref = os.path.join(ROOT, "oracle", "_ref", "libopal_ref.so")
    if os.path.exists(ref):
        return OpalCLibrary(ref), "reference"
    port = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(port):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return OpalCLibrary(port), "port"


def cpu_search(lib, sm, query, shards, search_type=OPAL_SEARCH_SCORE_END):
    """One search of `query` against all shards, one host thread per shard (the reference is
    re-entrant; ctypes releases the GIL). Returns wall seconds."""
    results = [new_results(len(s)) for s in shards]
    rcs = [0] * len(shards)

    def work(k):
        rcs[k], _ = lib.search_database(query, shards[k], GAP_OPEN, GAP_EXT, sm.flat(), sm.alphabet_length, results[k],
                                        search_type, MODES[MODE], OPAL_OVERFLOW_BUCKETS)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(shards))]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    assert all(rc == 0 for rc in rcs), rcs
    return dt


def split_for_threads(db, nthreads, stride=1):
    """Contiguous residue-balanced shards of every `stride`-th sequence of the length-sorted DB."""
    order = np.argsort(db.lengths, kind="stable")[::stride]
    lens = db.lengths[order].astype(np.int64)
    bounds = np.searchsorted(np.cumsum(lens), np.linspace(0, lens.sum(), nthreads + 1)[1:-1])
    parts = np.split(order, bounds)
    return [db.subset(p) for p in parts if len(p)], int(lens.sum())


def run_reference_arm(args, rank):
    if rank != 0:
        return
    sm, query, db, desc, scaling = make_workload(args.workload, 0, 1)
    lib, kind = cpu_library()
    cores = os.cpu_count() or 1
    # bounded sample: ~2 G cells per step keeps K steps within a couple of minutes on any host
    stride = max(1, int(len(query) * db.total_residues / 2.5e9))
    shards, residues = split_for_threads(db, cores, stride)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_search(lib, sm, query, shards)
    secs = sum(cpu_search(lib, sm, query, shards) for _ in range(args.steps))
    cells = len(query) * residues * args.steps
    value = cells / 1e9 / secs
    sample = f"every {stride}-th sequence of the length-sorted DB ({residues} residues), {len(shards)} threads, per step"
    line = {"impl": "reference", "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "int8/int16 (AVX2)" if kind == "reference" else "int64",
            "data": "synthetic", "config": {"workload": desc, "mode": MODE, "search": "score+end", "query_length": int(len(query))},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": len(shards), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def run_b200_arm(args, rank, local_rank, world):
    import torch
    from opal_b200.handle import OpalB200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; opal-b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    os.environ["OPAL_B200_DEVICE"] = str(local_rank)  # device of the drop-in entry points (they take no device argument)
    # one process per GPU on one host: the library's host pool (packing, result scatter) gets this rank's share of the cores
    os.environ.setdefault("OPAL_B200_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // max(world, 1))))
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    eng = OpalB200()
    sm, query, db, desc, scaling = make_workload(args.workload, rank, world)
    mat, A, Q = sm.flat(), sm.alphabet_length, int(len(query))
    handle = eng.create_db(db, local_rank)
    cells_rank = Q * db.total_residues
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_resident():
        flush.zero_()
        torch.cuda.synchronize()
        rc, sc, eq, et, ms = handle.search(query, GAP_OPEN, GAP_EXT, mat, A, OPAL_SEARCH_SCORE_END, MODE)
        if rc != 0:
            raise SystemExit(f"search failed rc={rc}: {eng.last_error()}")
        return ms, sc

    # ---- device-timed value (database resident)
    for _ in range(args.warmup):
        step_resident()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    wall0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(args.steps):
        ms, sc = step_resident()
        dev_ms += ms
        launches += handle.last_stats()["kernel_launches"]
    barrier()
    wall_resident = time.perf_counter() - wall0
    clk = clocks.stop()
    stats = handle.last_stats()
    t_dev = max_over_ranks(dev_ms / 1e3)
    total_cells = sum_over_ranks(float(cells_rank)) * args.steps
    value = total_cells / 1e9 / t_dev

    # ---- end to end through the drop-in C ABI (host buffers in, OpalSearchResult records out)
    def step_e2e():
        res = new_results(len(db))
        rc, res = eng.search_database(query, db, GAP_OPEN, GAP_EXT, mat, A, res, OPAL_SEARCH_SCORE_END, MODES[MODE],
                                      OPAL_OVERFLOW_BUCKETS)
        if rc != 0:
            raise SystemExit(f"opalSearchDatabase failed rc={rc}: {eng.last_error()}")
        return res

    for _ in range(min(args.warmup, 3)):
        res = step_e2e()
    assert (res["score"] == sc).all(), "drop-in call and resident handle disagree"
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_cells / 1e9 / t_e2e
    # one upload block [offsets | pair offsets | fold offsets | lengths | max code | residues] + one argument block
    # [counters | matrix | query]; the folded stream of the longest targets is built on the device (at most 128 offsets go up)
    n_fold = max(min(len(db) & ~1, 128), 1)
    index_bytes = (8 * (len(db) + 1) + 8 * max((len(db) + 1) // 2, 1) + 8 * n_fold + 4 * (max(len(db), 1) + 1) + 255) // 256 * 256
    h2d = int(index_bytes + db.total_residues + 64 + 1024 + (4 * A * A + 255) // 256 * 256 + Q + 16)
    d2h = int(3 * 4 * len(db) + 4)

    # ---- many queries against the resident database: same metric, several queries in flight
    nq = 32
    rng = np.random.default_rng(7)
    batch = [query] + [datasets.mutate(query, 0.5, rng, sm)[:Q] for _ in range(nq - 1)] if args.workload == "config2" else \
            [query] + [datasets.random_residues(Q, rng, sm) for _ in range(3)]
    batch_cells = float(sum(len(x) for x in batch)) * db.total_residues
    rc, *_ = handle.search_batch(batch, GAP_OPEN, GAP_EXT, mat, A, OPAL_SEARCH_SCORE_END, MODE, in_flight=3)
    if rc != 0:
        raise SystemExit(f"batch search failed rc={rc}: {eng.last_error()}")
    barrier()
    t0 = time.perf_counter()
    rc, _, _, _, batch_ms = handle.search_batch(batch, GAP_OPEN, GAP_EXT, mat, A, OPAL_SEARCH_SCORE_END, MODE, in_flight=3)
    barrier()
    t_batch_wall = max_over_ranks(time.perf_counter() - t0)
    t_batch_dev = max_over_ranks(batch_ms / 1e3)
    batch_total = sum_over_ranks(batch_cells)
    multi_query = {"queries": len(batch), "in_flight": 3, "value": batch_total / 1e9 / t_batch_dev, "unit": "GCUPS",
                   "wall_value": batch_total / 1e9 / t_batch_wall,
                   "note": "opalb200_db_search_batch on the resident database: device-timed (CUDA events, first launch of the "
                           "batch to last kernel end) and wall-clock with host buffers in and out"}

    # ---- roofline of the dominant kernel: packed-DPX issue rate (measured live) and HBM streaming
    peak_gcups, instr_per_s, _ = eng.measure_dpx_peak(local_rank)
    per_gpu = value / world
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = float(json.load(f)["hbm_gbs"])
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback"
    hbm_gbs = db.total_residues * args.steps / (dev_ms / 1e3) / 1e9
    traffic = ncu_dram_bytes_per_launch() if args.workload == "config2" else None
    roofline = {
        "bound": "dpx (integer pipe)", "achieved": per_gpu, "peak": peak_gcups, "unit": "GCUPS", "frac": per_gpu / peak_gcups,
        "traffic": traffic,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the dominant (bulk) search_kernel launch, from the committed "
                        "ncu --set full capture profiles/r1_ncu_search_kernel.txt; algorithmic bytes per launch = residues of its targets",
        "peak_source": f"measured live: {instr_per_s / 1e12:.2f} T packed s16x2 thread-instr/s x 2 cells / 6 instr (SW)",
        "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                "peak_source": hbm_src, "note": "algorithmic bytes = 1 B per DB residue per query (1/Q B per cell)"},
    }

    line = None
    if rank == 0:
        line = {"metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "s16x2 (DPX), s32 re-run on overflow", "data": "synthetic",
                "config": {"workload": desc, "mode": MODE, "search": "score+end", "query_length": Q,
                           "db_sequences_per_gpu": len(db), "db_residues_per_gpu": db.total_residues,
                           "l2": "256 MiB flush buffer written between timed steps",
                           "geometry": {k: stats[k] for k in ("G", "R", "passes", "warps_per_partition", "groups", "folded")}},
                "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": t_e2e / args.steps * 1e3, "path": "opalSearchDatabase (pack + H2D + kernels + D2H + records)"},
                "gpu_launches": launches, "clocks": clk, "roofline": roofline, "multi_query": multi_query,
                "wall_ms_per_step_resident": wall_resident / args.steps * 1e3}

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        lib, kind = cpu_library()
        cores = os.cpu_count() or 1
        stride = max(1, int(Q * db.total_residues / 2.5e9))
        shards, residues = split_for_threads(db, cores, stride)
        cpu_search(lib, sm, query, shards)
        reps, secs = 0, 0.0
        while secs < 10.0 and reps < 200:
            secs += cpu_search(lib, sm, query, shards)
            reps += 1
        line["cpu_baseline"] = {"value": Q * residues * reps / 1e9 / secs, "unit": "GCUPS", "cores": len(shards), "kind": kind,
                                "sample": f"every {stride}-th sequence of the length-sorted DB ({residues} residues) x {reps} "
                                          f"searches, one std thread per contiguous residue-balanced shard"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    handle.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
