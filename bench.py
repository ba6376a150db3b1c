#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on a B200: batched setitem MB/s of raw input (config[1], C2) as the headline
line, with the second half of the metric — batched getitem decode GB/s (config[2], C3) and batched contains lookup
(config[3], C4) — as sub-records of the same JSON line, each with its own e2e / roofline / cpu_baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode all|setitem|getitem|lookup|shard]

One step of the headline = one pass of the hot path (PiXiuCtrl::setitem for every record, in order) over the whole
C2 corpus (10k synthetic HTML-like pages x <= 60 KB, URL keys, ~356 MB raw), starting from a freshly rotated window.
Prints ONE JSON line (rank 0).

 value : raw-input MB/s, inputs already resident in HBM (pixiu_setitem_batch_dev), CUDA events on the store's stream
 e2e   : the same through pixiu_setitem_batch with pinned HOST buffers (H2D of keys/values and D2H of rc/saved inside
         the timed region)
 roofline     : dominant kernel class, algorithmic bytes / CUDA-event time vs the measured HBM copy bandwidth
 cpu_baseline : the unmodified reference (oracle/_ref) on this box's host cores, bounded sample
 getitem_c2   : batched getitem decode of every C2 record just stored (GB/s of encoded + decoded bytes)
 getitem_c3   : C3 = 1 M records x 1 KB with deeply nested back references: batched getitem of every record
 lookup_c4    : C4 = 10 M URL keys x ~200 B values (~2 GB compressed): batched contains of 10 M keys, 90 % present
 sharded      : (N > 1 only) BASELINE config 5 style: ONE extended window sharded over the N GPUs, every batch replicated,
                ncclAllReduce(MAX) of the match lengths and ncclAllReduce(MIN) of the leftmost candidates issued inside
                libpixiu_b200.so -> strong scaling, with the bytes and the stream time of both collectives

N > 1: one process per GPU (torchrun).  The headline and the read-side sub-records partition the work by key (every
rank owns its records in its own store; the reference window is ~12.5 MB, a shard never needs another shard's text):
no data-path collective, weak scaling.  The `sharded` sub-record is the mode with a real exchange step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "setitem_raw_input_throughput"
UNIT = "MB/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region, through NVML in a thread at 2 Hz, the two queries
    alternating.  (Round 1 measured large costs for faster sampling; those runs were confounded by the arena's mapping
    stalls - profiles/README.md - and `nvidia-smi -lms 250` next to the run now costs nothing measurable.  The rate stays
    low anyway: the step is a few thousand launches with a host synchronisation every half millisecond.)"""

    def __init__(self, index, first_delay=0.1, period=0.5):
        self.index = index
        self.first_delay, self.period = first_delay, period
        self.sm, self.reasons, self.mx = [], set(), None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                k = 0
                self._stop.wait(self.first_delay)
                while not self._stop.is_set():
                    try:
                        if k % 2 == 0:
                            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        else:
                            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                            for name, bit in bits.items():
                                if r & bit:
                                    self.reasons.add(name)
                    except Exception:
                        pass
                    k += 1
                    self._stop.wait(self.period)

            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
        except Exception as e:
            self.err = repr(e)
            self.t = None

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def gen_corpus(pages, seed):
    from pixiu_b200 import synth

    return synth.gen_html_pages(pages, seed=seed)


def _ref_worker(job):
    """one process = one instance of the reference (it is single-threaded and not re-entrant) on its own pages"""
    pages, seed = job
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    from pixiu_b200 import synth

    kd, ko, vd, vo = synth.gen_html_pages(pages, seed=seed)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    raw = int(ko[-1] + vo[-1])
    if po.ref_available():
        ref = po.Ref()
        r = ref.setitem_batch(keys, vals)
        sec, enc = float(r["seconds"]), int(r["enc_len"].sum())
        ref.close()
        kind = "reference"
    else:
        w = po.OracleWindow(strict251=True)
        t0 = time.perf_counter()
        enc = 0
        for k, v in zip(keys, vals):
            enc += len(w.encode(po.make_doc(k, v)))
        sec, kind = time.perf_counter() - t0, "port"
    return raw, sec, enc, kind


def ref_all_cores(pages_per_proc, procs):
    """the reference on every host core: `procs` independent instances (partition by key, like the multi-GPU mode),
    each ingesting `pages_per_proc` pages of the workload's generator; aggregate = total raw bytes / slowest instance"""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, [(pages_per_proc, 1000 + i) for i in range(procs)])
    raw = sum(r[0] for r in res)
    sec = max(r[1] for r in res)
    return {"value": raw / sec / 1e6, "unit": UNIT, "cores": procs, "kind": res[0][3],
            "sample": f"{procs} independent single-threaded instances (the reference is not re-entrant), {pages_per_proc} pages "
                      f"of the workload's generator each ({raw / 1e6:.1f} MB raw in total); aggregate = total / slowest instance",
            "seconds": sec, "stored_over_raw_on_sample": sum(r[2] for r in res) / raw}




def _pool_map(fn, jobs, procs):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        return pool.map(fn, jobs)


# ---- CPU legs of the read side: the reference's getitem / contains (PiXiuCtrl.cpp:55-61) on a bounded sample ----
def _ref_getitem_worker(job):
    """one reference instance: ingest `records` C3-style records (untimed), then time getitem + drain of every record
    (PXSGen::operator(), PiXiuStr.h:129-198).  The reference decoder has bugs B1/B2: its output is NOT
    checked, only timed."""
    records, seed = job
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    from pixiu_b200 import synth

    kd, ko, vd, vo = synth.gen_nested(records, seed=seed)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    if po.ref_available():
        ref = po.Ref()
        r = ref.setitem_batch(keys, vals)
        g = ref.getitem_batch(keys)
        ref.close()
        return int(r["enc_len"].sum()), int(g["bytes"]), float(g["seconds"]), "reference"
    st = po.OracleStore(strict251=True)
    enc = 0
    for k, v in zip(keys, vals):
        st.setitem(k, v)
    t0 = time.perf_counter()
    dec = sum(len(st.getitem(k)) for k in keys)
    return enc, dec, time.perf_counter() - t0, "port"


def cpu_getitem_c3(records, procs):
    res = _pool_map(_ref_getitem_worker, [(records, 1003 + i) for i in range(procs)], procs)
    alg = sum(r[0] + r[1] for r in res)
    sec = max(r[2] for r in res)
    return {"value": alg / sec / 1e9, "unit": "GB/s", "cores": procs, "kind": res[0][3], "seconds": sec,
            "decoded_mb_s": sum(r[1] for r in res) / sec / 1e6,
            "sample": f"{procs} independent single-threaded instances, each ingests {records} records of the C3 generator "
                      f"(untimed) and then drains getitem of every one of them (timed; output unchecked: reference "
                      f"decoder bugs B1/B2); aggregate = (encoded + decoded bytes) / slowest instance"}


def _ref_contains_worker(job):
    keys_n, seed = job
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as po
    from pixiu_b200 import synth

    kd, ko, vd, vo = synth.gen_urls_kv(keys_n, seed=seed, val_words=20)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    rng = np.random.default_rng(seed)
    n_abs = keys_n // 10
    q = [keys[i] for i in rng.integers(0, keys_n, size=keys_n - n_abs)] + [b"http://absent.qq.com/a/%d.htm" % i for i in range(n_abs)]
    q = [q[i] for i in rng.permutation(len(q))]
    if po.ref_available():
        ref = po.Ref()
        ref.setitem_batch(keys, vals)
        sec, found = 0.0, 0
        for _ in range(REF_CONTAINS_PASSES):     # (one pass is ~0.1 s: repeat it so that the timed part is seconds)
            c = ref.contains_batch(q)
            sec += float(c["seconds"])
            found += int(c["found"].sum())
        ref.close()
        return len(q) * REF_CONTAINS_PASSES, sec, found, "reference"
    st = po.OracleStore(strict251=True)
    for k, v in zip(keys, vals):
        st.setitem(k, v)
    t0 = time.perf_counter()
    f = sum(st.contains(k) for k in q)
    return len(q), time.perf_counter() - t0, f, "port"


REF_CONTAINS_PASSES = 20


def cpu_contains_c4(keys_n, procs):
    res = _pool_map(_ref_contains_worker, [(keys_n, 1004 + i) for i in range(procs)], procs)
    tot = sum(r[0] for r in res)
    sec = max(r[1] for r in res)
    return {"value": tot / sec / 1e6, "unit": "Mkeys/s", "cores": procs, "kind": res[0][3], "seconds": sec,
            "found_fraction": sum(r[2] for r in res) / tot,
            "sample": f"{procs} independent single-threaded instances, each ingests {keys_n} C4-style records (untimed) and "
                      f"then answers contains for {keys_n} keys, 90 % present / 10 % absent, random order, {REF_CONTAINS_PASSES} passes (timed); a "
                      f"{keys_n}-key tree is shallower than the 10 M-key one, which favours the CPU; aggregate = keys / slowest instance"}


def workload_config(args):
    """the workload both arms are quoted on (everything measured goes into `results`, so that the two lines carry the
    same `config`)"""
    return {"workload": f"C2: {args.pages} synthetic HTML-like pages x <=60KB (bytes 33..126), URL keys, batched setitem, per GPU",
            "records_per_gpu": args.pages, "window_policy": args.window,
            "l2": "inputs (~400 MB) and per-window scratch (~0.9 GB) exceed the 126 MB L2; no explicit flush",
            "partitioning": "by key across ranks, no collective"}


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """the reference's own CPU implementation (unmodified, compiled into oracle/_ref) on ALL host cores: one
    single-threaded instance per core on its own pages (the reference is not re-entrant; partition by key is also how
    the GPU arm scales).  One step = every instance ingests --ref-pages pages; value = total raw bytes / slowest instance.
    The read-side legs (getitem C3, contains C4) are timed once each and reported as sub-records."""
    if rank != 0:
        return
    procs = args.ref_procs or (os.cpu_count() or 1)
    n_warm = args.warmup   # (a pass is ~5 s of CPU work on every core)
    for _ in range(n_warm):
        ref_all_cores(args.ref_pages, procs)
    runs = [ref_all_cores(args.ref_pages, procs) for _ in range(args.steps)]
    val = float(np.mean([r["value"] for r in runs]))
    sec = float(np.mean([r["seconds"] for r in runs]))
    cb = dict(runs[-1])
    cb["value"] = val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "results": {"sample": f"the reference is timed on a bounded sample of the workload: {procs} instances x "
                                  f"{args.ref_pages} pages of its generator per step"},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.mode == "all" and not args.no_read_side:
        g = cpu_getitem_c3(args.ref_records, procs)
        line["getitem_c3"] = {"metric": "getitem_decode_throughput", "value": g["value"], "unit": "GB/s", "cpu_baseline": g,
                              "e2e": {"value": g["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        k = cpu_contains_c4(args.ref_keys, procs)
        line["lookup_c4"] = {"metric": "contains_lookup_throughput", "value": k["value"], "unit": "Mkeys/s", "cpu_baseline": k,
                             "e2e": {"value": k["value"], "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
class Env:
    """process-wide plumbing of the GPU arm: device, optional NCCL process group, barrier, timed loop"""

    def __init__(self, rank, world, local_rank):
        import torch

        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        self.torch, self.rank, self.world, self.local_rank = torch, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def ext_stream(self, c):
        return self.torch.cuda.ExternalStream(c.stream(), device=self.dev)

    def timed(self, ext, fn, steps):
        """K steps bracketed by barrier + synchronize, CUDA events on the store's stream, MAX over ranks"""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        self.barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall * 1e3], dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def reference_full_corpus_ratio(pages):
    """stored/raw of the UNMODIFIED reference over the full C2 corpus, from the committed fixture
    (tests/golden/c2_full_enc_len.npz, made by tests/golden/make_c2_full.py); None when it does not apply"""
    p = os.path.join(ROOT, "tests", "golden", "c2_full_enc_len.npz")
    if pages != 10000 or not os.path.exists(p):
        return None
    g = np.load(p)
    return float(g["enc_len"].astype(np.int64).sum()) / float(g["raw_bytes"])


def bench_setitem(args, env):
    """headline: C2 batched setitem; returns the fields of the main JSON line (rank 0) and the corpus for getitem_c2"""
    torch = env.torch
    from pixiu_b200 import ctrl, synth

    rank, world, dev = env.rank, env.world, env.dev
    kd, ko, vd, vo = gen_corpus(args.pages, 2 + rank)
    n = len(ko) - 1
    raw = int(ko[-1] + vo[-1])
    t_kd, t_ko, t_vd, t_vo = (torch.from_numpy(a) for a in (kd, ko, vd, vo))
    d_kd, d_ko, d_vd, d_vo = (t.to(dev) for t in (t_kd, t_ko, t_vd, t_vo))
    p_kd, p_ko, p_vd, p_vo = (t.pin_memory() for t in (t_kd, t_ko, t_vd, t_vo))
    policy = {"reference": ctrl.ROTATE_REFERENCE, "bytes": ctrl.ROTATE_BYTES, "records": ctrl.ROTATE_RECORDS}[args.window]
    c = ctrl.PiXiuCtrl(device=env.local_rank, rotate_policy=policy, window_bytes=args.window_bytes)
    ext = env.ext_stream(c)
    # The store grows by ~0.82 x raw per step.  The compressed arena maps HBM ahead of use on a helper thread (on a freshly
    # booted box one cuMemCreate / cuMemSetAccess of 256 MiB was measured at 20 - 200 ms, profiles/README.md), so the steps
    # do not wait for it; --reserve additionally passes the run's size as a capacity hint (pixiu_reserve)
    if args.reserve:
        passes = 2 * args.steps + args.warmup + min(args.warmup, 1) + 2
        c.reserve(int(0.86 * raw * passes) + (512 << 20))

    def step_dev():
        c.rotate()
        return c.setitem_batch_dev(d_kd.data_ptr(), d_ko.data_ptr(), d_vd.data_ptr(), d_vo.data_ptr(), n)

    def step_host():
        c.rotate()
        rc = np.zeros(n, dtype=np.int32)
        saved = np.zeros(n, dtype=np.int32)
        r = c._L.pixiu_setitem_batch(c._h, n, p_kd.data_ptr(), p_ko.data_ptr(), p_vd.data_ptr(), p_vo.data_ptr(),
                                     rc.ctypes.data_as(ctrl._i32p), saved.ctypes.data_as(ctrl._i32p))
        c._check(r)
        return rc, saved

    for _ in range(args.warmup):
        step_dev()
    # only rank 0 (whose line is printed) samples its GPU
    sampler = ClockSampler(env.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    st0 = c.stats()
    dev_ms, dev_wall_ms = env.timed(ext, step_dev, args.steps)
    st1 = c.stats()
    clocks = sampler.stop() if sampler else None
    launches_per_step = (st1.kernel_launches - st0.kernel_launches) // max(args.steps, 1)
    ratio = (st1.encoded_bytes - st0.encoded_bytes) / max(st1.raw_bytes - st0.raw_bytes, 1)
    chunks_per_step = (st1.chunks - st0.chunks) / max(args.steps, 1)

    for _ in range(min(args.warmup, 1)):
        step_host()
    e2e_ms, e2e_wall_ms = env.timed(ext, step_host, args.steps)

    # ---- roofline: per-kernel-class CUDA-event timing of one extra (untimed) step ----
    c.profile_enable(True)
    step_dev()
    prof = c.profile()
    c.profile_enable(False)
    peak, peak_kind = measured_peak()
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    top = max((k for k in prof if prof[k]["launches"]), key=lambda k: prof[k]["ms"])
    tp = prof[top]
    ach = tp["bytes"] / 1e9 / (tp["ms"] / 1e3) if tp["ms"] else 0.0
    traffic = None
    if top in NCU_TRAFFIC_OVER_ALGORITHMIC and tp["launches"]:
        traffic = NCU_TRAFFIC_OVER_ALGORITHMIC[top][0] * tp["bytes"] / tp["launches"]
    roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic,
                "traffic_source": (NCU_TRAFFIC_OVER_ALGORITHMIC[top][1] + " (ratio of a committed ncu capture x this run's "
                                   "algorithmic bytes per launch; not measured in this run)") if traffic else None,
                "algorithmic_bytes_per_launch": tp["bytes"] / max(tp["launches"], 1), "launches": tp["launches"],
                "avg_launch_ms": tp["ms"] / max(tp["launches"], 1), "share_of_step": tp["ms"] / tot_ms,
                "classes_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}}

    # ---- getitem_c2: batched getitem decode of everything just stored (device output), every record checked ----
    getitem = None
    try:
        getitem = getitem_measure(args, env, c, (kd, ko, vd, vo), "c2", steps=max(3, min(args.steps, 10)), cpu=None)
    except Exception as e:  # the headline metric stands on its own
        getitem = {"error": repr(e)}
    st_end = c.stats()
    c.free_prop()
    del d_kd, d_ko, d_vd, d_vo

    # ---- the reference-exact ratio: strict251 reproduces the reference's bytes (bug B1 included) ----
    strict_ratio = None
    try:
        cs = ctrl.PiXiuCtrl(device=env.local_rank, rotate_policy=policy, window_bytes=args.window_bytes, strict251=True)
        cs.setitem_batch((kd, ko), (vd, vo))
        ss = cs.stats()
        strict_ratio = ss.encoded_bytes / max(ss.raw_bytes, 1)
        cs.free_prop()
    except Exception as e:
        strict_ratio = repr(e)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = args.ref_procs or (os.cpu_count() or 1)
        cpu = ref_all_cores(args.ref_pages, procs)
        one = _ref_worker((args.ref_pages, 2))          # one instance alone on the first pages of the workload itself
        cpu["single_instance"] = {"value": one[0] / one[1] / 1e6, "unit": UNIT, "cores": 1,
                                  "sample": f"first {args.ref_pages} pages of the workload ({one[0] / 1e6:.1f} MB raw)",
                                  "stored_over_raw_on_sample": one[2] / one[0]}
        cpu["host_cores"] = os.cpu_count()

    total_raw = raw * world
    line = {
        "metric": METRIC, "value": total_raw * args.steps / (dev_ms / 1e3) / 1e6, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "results": {"raw_bytes_per_gpu": raw, "windows_per_step": chunks_per_step, "stored_over_raw": ratio,
                   "stored_over_raw_strict251": strict_ratio,
                   "stored_over_raw_reference_full_corpus": reference_full_corpus_ratio(args.pages) if rank == 0 else None,
                   "ratio_note": "default mode writes a run of exactly 251 bytes in the 8-byte form (+2 B per such run; the "
                                 "reference's 6-byte form FB FB.. is undecodable, bug B1); strict251 reproduces the reference's "
                                 "bytes; *_reference_full_corpus is the unmodified reference over the same 10,000 pages "
                                 "(tests/golden/c2_full_enc_len.npz)",
                   "steps_note": "every step re-inserts the same keys: after the first pass each record takes the index's "
                                 "REPLACE path and tombstones its previous copy; nothing is reclaimed between steps (the "
                                 f"store grows by ~{(st1.encoded_bytes - st0.encoded_bytes) / max(args.steps, 1) / 1e6:.0f} MB per step)"},
        "wall_ms_per_step": dev_wall_ms / args.steps,
        "e2e": {"value": total_raw * args.steps / (e2e_ms / 1e3) / 1e6, "unit": UNIT,
                "h2d_bytes_per_step": int(kd.nbytes + ko.nbytes + vd.nbytes + vo.nbytes), "d2h_bytes_per_step": int(8 * n),
                "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "cpu_baseline": cpu, "getitem_c2": getitem,
        "index_memory": {"key_arena_bytes": int(st_end.index_key_arena_bytes), "host_bytes": int(st_end.index_host_bytes),
                         "device_bytes": int(st_end.index_device_bytes), "table_device_bytes": int(st_end.table_device_bytes),
                         "note": "held beside the compressed store (pixiu_stats.index_*): the index keeps esc(k) 251 0 of every "
                                 "leaf instead of decoding the record to verify a key"},
        "clocks": clocks,
    }
    return line


# DRAM traffic over algorithmic bytes of the dominant kernels, from the committed `ncu --set full` captures under profiles/
NCU_TRAFFIC_OVER_ALGORITHMIC = {
    # k_rs_onesweep: 150.06 MB read + 100.93 MB written per launch of 12 M keys vs n x 24 B (L2 absorbs part of the scatter)
    "sort_pass": ((150.06e6 + 100.93e6) / (12e6 * 24), "profiles/r01_onesweep_12Mkeys_ncu_raw.csv"),
}
NCU_DECODE_TRAFFIC = {
    # k_decode_literals + k_decode_copies, dram__bytes_read.sum + dram__bytes_write.sum of one decode call over its
    # algorithmic bytes.  c2: 3,000 pages, (92.0 + 83.4) + (142.8 + 69.3) MB for 196.5 MB; c3: 400,000 records,
    # (71.0 + 409.6) + (332.7 + 360.3) MB for 422.8 MB (the literals kernel writes every byte once - reference bytes as
    # zero - the copy kernel writes the reference bytes again and reads their sources and the 8-byte pieces)
    "c2": ((92.02 + 83.41 + 142.78 + 69.25) / 196.5,
           "profiles/r02_decode_c2_3000pages_ncu_raw.csv (ratio of a committed ncu capture x this run's algorithmic bytes; "
           "not measured in this run)"),
    "c3": ((70.96 + 409.62 + 332.74 + 360.32) / 422.8,
           "profiles/r02_decode_c3_400krecords_ncu_raw.csv (ratio of a committed ncu capture x this run's algorithmic bytes; "
           "not measured in this run)"),
}


def getitem_measure(args, env, c, corpus, workload, steps, cpu):
    """batched getitem of EVERY record of `corpus` already stored in `c`; returns a sub-record
      value : GB/s of (encoded bytes read + decoded bytes written), whole pixiu_getitem_batch_dev call timed with CUDA
              events on the store's stream (keys are host inputs, the decoded output stays in HBM)
      e2e   : pixiu_getitem_batch into a pinned HOST buffer (key H2D, decoded bytes D2H inside the timed region)
      roofline : the decode kernels alone, algorithmic bytes / event time"""
    torch = env.torch
    from pixiu_b200 import synth

    kd, ko, vd, vo = corpus
    n = len(ko) - 1
    st = c.stats()
    ext = env.ext_stream(c)
    # decoded docs of the LIVE copies of the corpus' records: key + value + 4 terminator bytes each (escape-free data)
    dec_bytes = int(ko[-1] + vo[-1] + 4 * n)
    cap = dec_bytes + 64
    out = torch.empty(cap, dtype=torch.uint8, device=env.dev)
    hout = torch.empty(cap, dtype=torch.uint8).pin_memory()
    hout_np = hout.numpy()

    def step_dev():
        return c.getitem_batch_dev((kd, ko), out.data_ptr(), cap)

    def step_host():
        return c.getitem_batch((kd, ko), out=hout_np)

    for _ in range(max(args.warmup, 3) if workload != "c2" else 3):
        off, found = step_dev()
    assert found.all()
    sampler = ClockSampler(env.local_rank, first_delay=0.002, period=0.01) if env.rank == 0 else None
    if sampler:
        sampler.start()
    l0 = c.stats().kernel_launches
    dev_ms, dev_wall = env.timed(ext, step_dev, steps)
    launches = c.stats().kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    step_host()
    e2e_ms, e2e_wall = env.timed(ext, step_host, steps)
    c.profile_enable(True)
    step_dev()
    pd = c.profile()["decode"]
    c.profile_enable(False)
    s = c.stats()
    peak, peak_kind = measured_peak()
    alg = float(pd["bytes"])                      # encoded bytes of the decoded ranges + decoded bytes
    ach = alg / 1e9 / (pd["ms"] / 1e3)
    ncu = NCU_DECODE_TRAFFIC.get(workload)
    # bit-exact round trip of EVERY record against the inputs (escape-free synthetic data)
    hb = out[: int(off[-1])].cpu().numpy()
    klen, vlen = np.diff(ko), np.diff(vo)
    ok = bool(np.array_equal(np.diff(off), klen + vlen + 4)
              and np.array_equal(synth.ragged_gather(hb, off[:-1], klen), kd)
              and np.array_equal(synth.ragged_gather(hb, off[:-1] + klen + 2, vlen), vd))
    if not ok:
        raise SystemExit(f"getitem bench ({workload}): decoded records differ from the inputs")
    world = env.world
    return {"metric": "getitem_decode_throughput", "value": alg * world * steps / (dev_ms / 1e3) / 1e9, "unit": "GB/s",
            "n_gpus": world, "steps": steps, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak",
            "config": {"workload": workload, "records_per_gpu": n, "decoded_bytes_per_gpu": dec_bytes,
                       "algorithmic_bytes_per_gpu": alg, "chunks": int(st.chunks),
                       "bytes_counted": "encoded read + decoded written (SURVEY 8d)",
                       "l2": "decoded output (>= 350 MB) exceeds the 126 MB L2; no explicit flush"},
            "wall_ms_per_step": dev_wall / steps,
            "e2e": {"value": alg * world * steps / (e2e_ms / 1e3) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": int(kd.nbytes + ko.nbytes), "d2h_bytes_per_step": int(off[-1]) + 9 * n,
                    "ms_per_step": e2e_ms / steps, "wall_ms_per_step": e2e_wall / steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_decode_literals + k_decode_copies", "achieved": ach, "peak": peak,
                         "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak,
                         "traffic": (ncu[0] * alg) if ncu else None, "traffic_source": ncu[1] if ncu else None,
                         "algorithmic_bytes_per_launch": alg, "avg_launch_ms": pd["ms"], "launches": pd["launches"],
                         "lookup_ms": s.last_lookup_gpu_ms},
            "roundtrip_all_records_ok": ok, "clocks": clocks, "cpu_baseline": cpu}


def bench_getitem(args, env, workload):
    """BASELINE config[2] (C3): 1 M records x 1 KB with deeply nested back references (or the C2 pages): setitem of the
    corpus (untimed), then getitem_measure; the CPU leg runs the reference's getitem on a bounded sample (rank 0, N=1)"""
    from pixiu_b200 import ctrl, synth

    if workload == "c2":
        corpus = gen_corpus(args.pages, 2 + env.rank)
        wl = f"C2 decode: {args.pages} synthetic HTML-like pages x <=60KB per GPU, batched getitem of every record"
    else:
        corpus = synth.gen_nested(args.records, seed=3 + env.rank)
        wl = (f"C3: {args.records} records x 1 KB per GPU, record r = record r-1 with one swept byte mutated "
              f"(deeply nested back references) + self-periodic runs, batched getitem of every record")
    c = ctrl.PiXiuCtrl(device=env.local_rank, rotate_policy=ctrl.ROTATE_REFERENCE)
    c.setitem_batch((corpus[0], corpus[1]), (corpus[2], corpus[3]))
    cpu = None
    if env.rank == 0 and env.world == 1 and not args.no_cpu and workload == "c3":
        cpu = cpu_getitem_c3(args.ref_records, args.ref_procs or (os.cpu_count() or 1))
    rec = getitem_measure(args, env, c, corpus, workload, steps=args.steps, cpu=cpu)
    rec["config"]["workload"] = wl
    st = c.stats()
    rec["config"]["encoded_bytes_per_gpu"] = int(st.encoded_bytes)
    rec["config"]["stored_over_raw"] = st.encoded_bytes / max(st.raw_bytes, 1)
    c.free_prop()
    return rec


def bench_lookup(args, env):
    """BASELINE config[3] (C4): batched contains over N URL keys with ~200-byte values (10 M keys = a ~2 GB compressed
    corpus); the query batch is N keys, 90 % present / 10 % absent, in random order.  One step = one batch.
      value : M keys/s, queries already packed in HBM (pixiu_contains_batch_dev), CUDA events on the store's stream
      e2e   : pixiu_contains_batch with HOST buffers (query H2D and found[] D2H inside the timed region)
      roofline : k_query_len/write + k_lookup, algorithmic bytes per query = q_len + depth x 7 + q_len"""
    torch = env.torch
    from pixiu_b200 import ctrl, synth

    rank, world, dev = env.rank, env.world, env.dev
    n = args.keys
    kd, ko, vd, vo = synth.gen_urls_kv(n, seed=4 + rank, val_words=20)
    c = ctrl.PiXiuCtrl(device=env.local_rank, rotate_policy=ctrl.ROTATE_REFERENCE)
    t0 = time.perf_counter()
    c.setitem_batch((kd, ko), (vd, vo))
    set_s = time.perf_counter() - t0
    st = c.stats()
    del vd
    # queries
    rng = np.random.default_rng(40 + rank)
    n_abs = n // 10
    n_pre = n - n_abs
    pick = rng.integers(0, n, size=n_pre)
    ad, ao = synth.pack([b"http://absent.qq.com/a/%d.htm" % i for i in range(n_abs)])
    klen = np.diff(ko)
    all_d = np.concatenate([kd, ad])
    starts = np.concatenate([ko[:-1][pick], ko[-1] + ao[:-1]])
    lens = np.concatenate([klen[pick], np.diff(ao)])
    perm = rng.permutation(n)
    qd = synth.ragged_gather(all_d, starts[perm], lens[perm])
    qo = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens[perm], out=qo[1:])
    expect = perm < n_pre
    d_qd, d_qo = torch.from_numpy(qd).to(dev), torch.from_numpy(qo).to(dev)
    d_found = torch.zeros(n, dtype=torch.uint8, device=dev)
    p_qd, p_qo = torch.from_numpy(qd).pin_memory(), torch.from_numpy(qo).pin_memory()
    found_h = np.zeros(n, dtype=np.uint8)
    ext = env.ext_stream(c)

    def step_dev():
        c.contains_batch_dev(d_qd.data_ptr(), d_qo.data_ptr(), n, d_found.data_ptr())

    def step_host():
        c._check(c._L.pixiu_contains_batch(c._h, n, p_qd.data_ptr(), p_qo.data_ptr(), found_h.ctypes.data_as(ctrl._u8p)))

    for _ in range(max(args.warmup, 3)):
        step_dev()
    if not np.array_equal(d_found.cpu().numpy().astype(bool), expect):
        raise SystemExit("lookup bench: found[] differs from the expected presence of the queries")
    sampler = ClockSampler(env.local_rank, first_delay=0.002, period=0.01) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = c.stats().kernel_launches
    dev_ms, dev_wall = env.timed(ext, step_dev, args.steps)
    launches = c.stats().kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    step_host()
    if not np.array_equal(found_h.astype(bool), expect):
        raise SystemExit("lookup bench: host-path found[] differs")
    e2e_ms, e2e_wall = env.timed(ext, step_host, args.steps)
    c.profile_enable(True)
    step_dev()
    pl = c.profile()["lookup"]
    c.profile_enable(False)
    # mean depth of the walks on a sample (host index)
    samp = np.arange(0, n, max(n // 20000, 1))
    sd = synth.ragged_gather(qd, qo[:-1][samp], np.diff(qo)[samp])
    so = np.zeros(len(samp) + 1, dtype=np.int64)
    np.cumsum(np.diff(qo)[samp], out=so[1:])
    depth = c.debug_index_depth((sd, so))
    mean_depth = float(depth.mean())
    alg = pl["bytes"] + n * mean_depth * 7.0
    peak, peak_kind = measured_peak()
    ach = alg / 1e9 / (pl["ms"] / 1e3)
    st2 = c.stats()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_contains_c4(args.ref_keys, args.ref_procs or (os.cpu_count() or 1))
    rec = {"metric": "contains_lookup_throughput", "value": n * world * args.steps / (dev_ms / 1e3) / 1e6, "unit": "Mkeys/s",
           "n_gpus": world, "steps": args.steps, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "config": {"workload": f"C4: {n} URL keys x ~200 B values per GPU stored ({st.chunks} chunks, {st.encoded_bytes / 1e9:.2f} GB "
                                  f"compressed), batched contains of {n} keys, 90 % present / 10 % absent, random order",
                      "keys_per_gpu": n, "query_bytes": int(qd.nbytes), "mean_walk_depth": mean_depth,
                      "max_walk_depth_in_sample": int(depth.max()), "stored_over_raw": st.encoded_bytes / max(st.raw_bytes, 1),
                      "index_key_arena_bytes": int(st2.index_key_arena_bytes), "index_device_bytes": int(st2.index_device_bytes),
                      "index_host_bytes": int(st2.index_host_bytes),
                      "stored_plus_index_over_raw": (st.encoded_bytes + st2.index_device_bytes) / max(st.raw_bytes, 1),
                      "setitem_seconds_incl_index": set_s, "setitem_mb_s": st.raw_bytes / set_s / 1e6,
                      "setitem_gpu_ms": st.last_setitem_gpu_ms,
                      "setitem_note": "one cold call of the process (first allocations included); wall = GPU encode + CritBit insert",
                      "l2": "index (nodes + key arena) and queries exceed the 126 MB L2 at 10 M keys; no explicit flush"},
           "wall_ms_per_step": dev_wall / args.steps,
           "e2e": {"value": n * world * args.steps / (e2e_ms / 1e3) / 1e6, "unit": "Mkeys/s",
                   "h2d_bytes_per_step": int(qd.nbytes + qo.nbytes), "d2h_bytes_per_step": int(n),
                   "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall / args.steps},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "k_query_len + scan + k_query_write + k_lookup", "achieved": ach, "peak": peak,
                        "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": 433.7 * n,
                        "traffic_source": "profiles/r01_lookup_3Mkeys_ncu_raw.csv: k_lookup (unchanged since) read 1.287 GB and "
                                          "wrote 14.7 MB of DRAM for 3 M queries = 433.7 B per query, x this run's queries; "
                                          "not measured in this run",
                        "algorithmic_bytes_per_launch": alg, "avg_launch_ms": pl["ms"], "launches": pl["launches"],
                        "sector_granular_gbs": (pl["bytes"] + n * mean_depth * 32.0) / 1e9 / (pl["ms"] / 1e3),
                        "bytes_model": "2 x escaped query bytes + depth x 7 B per query (SURVEY 8d); sector-granular: depth x 32 B"},
           "found_matches_expected": True, "clocks": clocks, "cpu_baseline": cpu}
    c.free_prop()
    return rec


def bench_sharded(args, env):
    """BASELINE config 5 style: ONE extended window sharded by record over the GPUs; every batch goes to all ranks;
    match lengths are MAX-reduced and leftmost candidates MIN-reduced by ncclAllReduce inside libpixiu_b200.so
    (pixiu_mg_setitem_batch).  Total work is fixed as N grows: strong scaling."""
    torch = env.torch
    from pixiu_b200 import ctrl, multigpu, shard, synth

    rank, world = env.rank, env.world
    # The same corpus on every rank.  Large corpora (BASELINE config 5: 8 GB = ~200k pages) are generated in 8 parts,
    # each by one rank, and broadcast: the page generator is single-threaded Python/numpy (~1 min per GB).
    # The corpus is the same for every N (strong scaling).
    PARTS = 8
    if args.shard_pages >= 16 * PARTS:
        per = args.shard_pages // PARTS
        parts = [None] * PARTS
        for p in range(PARTS):
            if p % world == rank:
                a = synth.gen_html_pages(per if p < PARTS - 1 else args.shard_pages - per * (PARTS - 1), seed=500 + p)
                # (keys are only unique inside one generator call: tag them with the part)
                tag = np.frombuffer(b"/p%d" % p, dtype=np.uint8)
                klen = np.diff(a[1]) + len(tag)
                ko2 = np.zeros(len(klen) + 1, dtype=np.int64)
                np.cumsum(klen, out=ko2[1:])
                kd2 = np.empty(int(ko2[-1]), dtype=np.uint8)
                src_idx = synth.ragged_gather(np.arange(len(a[0]), dtype=np.int64), a[1][:-1], np.diff(a[1]))
                pos = np.repeat(ko2[:-1], np.diff(a[1])) + (np.arange(len(a[0])) - np.repeat(a[1][:-1], np.diff(a[1])))
                kd2[pos] = a[0][src_idx]
                for j in range(len(tag)):
                    kd2[ko2[1:] - len(tag) + j] = tag[j]
                parts[p] = (kd2, ko2, a[2], a[3])
        for p in range(PARTS if world > 1 else 0):
            owner = p % world
            hdr = torch.zeros(4, dtype=torch.int64, device=env.dev)
            if owner == rank:
                hdr = torch.tensor([len(x) for x in parts[p]], dtype=torch.int64, device=env.dev)
            env.dist.broadcast(hdr, src=owner)
            out = []
            for j, (n_el, dt) in enumerate(zip(hdr.tolist(), (torch.uint8, torch.int64, torch.uint8, torch.int64))):
                t = torch.from_numpy(parts[p][j]).to(env.dev) if owner == rank else torch.empty(n_el, dtype=dt, device=env.dev)
                env.dist.broadcast(t, src=owner)
                out.append(t.cpu().numpy())
                del t
            parts[p] = tuple(out)
        kd = np.concatenate([x[0] for x in parts])
        vd = np.concatenate([x[2] for x in parts])
        ko = np.concatenate([[0]] + [x[1][1:] + off for x, off in zip(parts, np.cumsum([0] + [int(x[1][-1]) for x in parts[:-1]]))]).astype(np.int64)
        vo = np.concatenate([[0]] + [x[3][1:] + off for x, off in zip(parts, np.cumsum([0] + [int(x[3][-1]) for x in parts[:-1]]))]).astype(np.int64)
        del parts
        torch.cuda.empty_cache()
    else:
        kd, ko, vd, vo = gen_corpus(args.shard_pages, 2)
    n = len(ko) - 1
    raw = int(ko[-1] + vo[-1])
    bp = args.batch_pages
    batches = []
    for a in range(0, n, bp):
        idx = np.arange(a, min(n, a + bp))
        batches.append((shard.take_packed(kd, ko, idx), shard.take_packed(vd, vo, idx)))

    def one_pass(k):
        c = ctrl.PiXiuCtrl(device=env.local_rank, rotate_policy=ctrl.ROTATE_BYTES, window_bytes=args.shard_window_bytes)
        if world > 1:
            multigpu.init_comm_torch(c)       # bootstrap only: the 128-byte NCCL id travels over the process group
        else:
            c.mg_comm_init(0, 1, c.mg_unique_id())
        # every pass starts from a NEW store: without the capacity hint its first batch would map the first 256 MiB of the
        # compressed arena inside the timed region (20 - 200 ms per cuMemCreate on a freshly booted box: the 2-GPU line
        # read 146 or 215 ms per pass depending on that alone)
        c.reserve(int(0.86 * raw) + (64 << 20))
        env.barrier()
        t0 = time.perf_counter()
        for kb, vb in batches:
            c.mg_setitem_batch(kb, vb)
        env.barrier()
        dt = time.perf_counter() - t0
        st, ms = c.stats(), c.mg_stats()
        out = dict(sec=dt, ratio=st.encoded_bytes / max(st.raw_bytes, 1), chunks=int(st.chunks), launches=int(st.kernel_launches),
                   max_ms=ms.max_reduce_ms, min_ms=ms.min_reduce_ms, max_bytes=int(ms.max_reduce_bytes),
                   min_bytes=int(ms.min_reduce_bytes), nccl=int(ms.nccl_version))
        c.free_prop()
        return out

    for k in range(min(args.warmup, 1)):
        one_pass(-1 - k)
    steps = max(1, min(args.steps, 3))
    res = [one_pass(k) for k in range(steps)]
    t = torch.tensor([sum(r["sec"] for r in res), sum(r["max_ms"] for r in res), sum(r["min_ms"] for r in res)],
                     dtype=torch.float64, device=env.dev)
    if env.dist is not None:
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
    sec = float(t[0]) / steps
    last = res[-1]
    return {"metric": METRIC, "value": raw / sec / 1e6, "unit": UNIT, "n_gpus": world, "steps": steps,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "timing": "wall clock around the pass, bracketed by barrier + synchronize, MAX over ranks (each batch ends "
                      "synchronised: rc/saved are returned to the host)",
            "config": {"workload": f"C5 style: {args.shard_pages} synthetic HTML-like pages ({raw / 1e6:.0f} MB raw), ONE window sharded "
                                   f"over {world} GPU(s), {args.shard_window_bytes} window bytes per GPU, batches of {bp} pages "
                                   f"replicated to all ranks",
                       "mode": "sharded window + ncclAllReduce(MAX) of M / ncclAllReduce(MIN) of leftmost candidates, issued by "
                               "libpixiu_b200.so on the store's stream",
                       "raw_bytes": raw, "stored_over_raw": last["ratio"], "chunks": last["chunks"], "batches_per_step": len(batches)},
            "gpu_launches": last["launches"],
            "collectives": {"per_step": 2 * len(batches) if world > 1 else 0, "nccl_version": last["nccl"],
                            "max_reduce_bytes_per_step": last["max_bytes"], "min_reduce_bytes_per_step": last["min_bytes"],
                            "max_reduce_ms_per_step": float(t[1]) / steps, "min_reduce_ms_per_step": float(t[2]) / steps,
                            "share_of_step": (float(t[1]) + float(t[2])) / steps / (sec * 1e3),
                            "note": "stream time between the events around each ncclAllReduce (MAX over ranks): includes "
                                    "waiting for the slowest rank's phase"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages", type=int, default=10000)
    ap.add_argument("--window", default="reference", choices=["reference", "bytes", "records"])
    ap.add_argument("--window-bytes", type=int, default=12_500_000)
    ap.add_argument("--ref-pages", type=int, default=100, help="bounded sample for the CPU reference leg: pages per instance")
    ap.add_argument("--ref-records", type=int, default=12000, help="CPU getitem leg: C3 records per instance")
    ap.add_argument("--ref-keys", type=int, default=30000, help="CPU contains leg: C4 keys per instance")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference instances run side by side (0 = one per host core)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reserve", action="store_true", help="pass the size of the run as a capacity hint (pixiu_reserve) before the steps")
    ap.add_argument("--no-read-side", action="store_true", help="mode all: skip the getitem_c3 / lookup_c4 sub-records")
    ap.add_argument("--mode", default="all", choices=["all", "setitem", "partition", "shard", "getitem", "lookup"],
                    help="all: headline setitem line + getitem_c3 / lookup_c4 (and, N > 1, sharded) sub-records; "
                         "setitem (= partition): the headline alone; getitem / lookup / shard: that record as the line")
    ap.add_argument("--batch-pages", type=int, default=1024)
    ap.add_argument("--shard-pages", type=int, default=3000, help="sharded mode: pages of the corpus (same on every rank)")
    ap.add_argument("--shard-window-bytes", type=int, default=64_000_000, help="sharded mode: window bytes per GPU")
    ap.add_argument("--workload", default="c3", choices=["c2", "c3"], help="--mode getitem: which corpus to decode")
    ap.add_argument("--keys", type=int, default=10_000_000, help="lookup: keys stored and queried per GPU")
    ap.add_argument("--records", type=int, default=1_000_000, help="getitem c3: records per GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    env = Env(rank, world, local_rank)
    base = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "vs_baseline": None, "dtype": "u8", "data": "synthetic"}
    if args.mode == "shard":
        line = dict(base, **bench_sharded(args, env))
    elif args.mode == "getitem":
        line = dict(base, **bench_getitem(args, env, args.workload))
    elif args.mode == "lookup":
        line = dict(base, **bench_lookup(args, env))
    else:
        line = bench_setitem(args, env)
        if args.mode == "all" and not args.no_read_side:
            for name, fn in (("getitem_c3", lambda: bench_getitem(args, env, "c3")), ("lookup_c4", lambda: bench_lookup(args, env))):
                try:
                    line[name] = fn()
                except SystemExit as e:   # a failed parity check of a sub-record must be visible, not fatal to the headline
                    line[name] = {"error": str(e)}
                except Exception as e:
                    line[name] = {"error": repr(e)}
        if args.mode == "all" and world > 1:
            try:
                line["sharded"] = bench_sharded(args, env)
            except Exception as e:
                line["sharded"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    env.close()


if __name__ == "__main__":
    main()
