#!/usr/bin/env python
"""bench.py — batched setitem (compress-on-insert) throughput on BASELINE.json config[1]:
10k synthetic HTML-like pages x <=60 KB ASCII with URL keys (~400 MB raw) on 1 x B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path (PiXiuCtrl::setitem for every record, in order) over the
whole corpus, starting from a freshly rotated window.  Prints ONE JSON line (rank 0).

 value : raw-input MB/s, inputs already resident in HBM (pixiu_setitem_batch_dev), CUDA events
 e2e   : same through pixiu_setitem_batch with pinned HOST buffers (H2D of keys/values and
         D2H of rc/saved inside the timed region)
 roofline     : dominant kernel class, algorithmic bytes / CUDA-event time vs measured HBM peak
 cpu_baseline : the unmodified reference (oracle/_ref) on this box's host cores, bounded sample
 getitem      : (extra) batched getitem decode of every record, GB/s of encoded+decoded bytes

N > 1: one process per GPU (torchrun); the corpus is partitioned by key, every rank ingests its
own pages into its own store (no data-path collective: the reference window is ~12.5 MB, a
shard never needs another shard's text) -> weak scaling.
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
    """SM clock / throttle reasons DURING the timed region, through NVML in a thread at 2 Hz.
    Sampling is not free for a launch-heavy step (measured on this pool, scratch/nvml_cost.py: spawning
    `nvidia-smi -lms 100` next to the run halves the throughput, NVML clock+reasons at 20 Hz costs 45 %,
    at 4 Hz 14 %, one query kind at 4 Hz nothing measurable), so the rate is kept low and the two queries
    alternate."""

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


# ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """the reference's own CPU implementation (unmodified, compiled into oracle/_ref) on ALL host cores: one
    single-threaded instance per core on its own pages (the reference is not re-entrant; partition by key is also how
    the GPU arm scales).  One step = every instance ingests --ref-pages pages; value = total raw bytes / slowest instance."""
    if rank != 0:
        return
    procs = args.ref_procs or (os.cpu_count() or 1)
    n_warm = max(args.warmup_ref, min(args.warmup, 1))  # one untimed pass at most: a pass is seconds of CPU work
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
            "config": {"workload": f"C2: {args.pages} synthetic HTML-like pages x <=60KB, URL keys, batched setitem (reference timed on a "
                                   f"bounded sample: {procs} instances x {args.ref_pages} pages per step)"},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch

    from pixiu_b200 import ctrl, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    kd, ko, vd, vo = gen_corpus(args.pages, 2 + rank)
    n = len(ko) - 1
    raw = int(ko[-1] + vo[-1])
    dev = torch.device("cuda", local_rank)
    t_kd, t_ko, t_vd, t_vo = (torch.from_numpy(a) for a in (kd, ko, vd, vo))
    d_kd, d_ko, d_vd, d_vo = (t.to(dev) for t in (t_kd, t_ko, t_vd, t_vo))
    p_kd, p_ko, p_vd, p_vo = (t.pin_memory() for t in (t_kd, t_ko, t_vd, t_vo))
    policy = {"reference": ctrl.ROTATE_REFERENCE, "bytes": ctrl.ROTATE_BYTES, "records": ctrl.ROTATE_RECORDS}[args.window]
    c = ctrl.PiXiuCtrl(device=local_rank, rotate_policy=policy, window_bytes=args.window_bytes)
    ext = torch.cuda.ExternalStream(c.stream(), device=dev)

    def step_dev():
        c.rotate()
        return c.setitem_batch_dev(d_kd.data_ptr(), d_ko.data_ptr(), d_vd.data_ptr(), d_vo.data_ptr(), n)

    def step_host():
        c.rotate()
        L = c._L
        rc = np.zeros(n, dtype=np.int32)
        saved = np.zeros(n, dtype=np.int32)
        r = L.pixiu_setitem_batch(c._h, n, p_kd.data_ptr(), p_ko.data_ptr(), p_vd.data_ptr(), p_vo.data_ptr(),
                                  rc.ctypes.data_as(ctrl._i32p), saved.ctypes.data_as(ctrl._i32p))
        c._check(r)
        return rc, saved

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        step_dev()
    # NVML queries take a driver-wide lock: with one sampler per rank the ranks of a launch-bound step slow each other
    # down, so only rank 0 (whose line is printed) samples its GPU
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    st0 = c.stats()
    dev_ms, dev_wall_ms = timed(step_dev, args.steps)
    st1 = c.stats()
    clocks = sampler.stop() if sampler else None
    launches_per_step = (st1.kernel_launches - st0.kernel_launches) // max(args.steps, 1)
    ratio = (st1.encoded_bytes - st0.encoded_bytes) / max(st1.raw_bytes - st0.raw_bytes, 1)
    chunks_per_step = (st1.chunks - st0.chunks) / max(args.steps, 1)

    for _ in range(min(args.warmup, 1)):
        step_host()
    e2e_ms, e2e_wall_ms = timed(step_host, args.steps)

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
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/
    # r01_onesweep_12Mkeys_ncu_raw.csv): 150.06 MB read + 100.93 MB written per launch of 12 M keys, i.e.
    # 0.871 x the algorithmic n x 24 B (L2 absorbs part of the scatter); scaled to this run's average launch.
    NCU_TRAFFIC_OVER_ALGORITHMIC = {"sort_pass": (150.06e6 + 100.93e6) / (12e6 * 24)}
    traffic = None
    if top in NCU_TRAFFIC_OVER_ALGORITHMIC and tp["launches"]:
        traffic = NCU_TRAFFIC_OVER_ALGORITHMIC[top] * tp["bytes"] / tp["launches"]
    roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": tp["bytes"] / max(tp["launches"], 1),
                "launches": tp["launches"],
                "avg_launch_ms": tp["ms"] / max(tp["launches"], 1), "share_of_step": tp["ms"] / tot_ms,
                "classes_ms": {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}}

    # ---- extra: batched getitem decode of everything just stored (device output) ----
    getitem = None
    try:
        out = torch.empty(int(raw + 8 * n + 64), dtype=torch.uint8, device=dev)
        c.getitem_batch_dev((kd, ko), out.data_ptr(), out.numel())  # warm
        torch.cuda.synchronize()
        tg0 = time.perf_counter()
        off, found = c.getitem_batch_dev((kd, ko), out.data_ptr(), out.numel())
        call_ms = (time.perf_counter() - tg0) * 1e3     # whole call: key upload, lookup, work list, decode (it returns synchronised)
        s = c.stats()
        c.profile_enable(True)
        c.getitem_batch_dev((kd, ko), out.data_ptr(), out.numel())
        pd = c.profile()["decode"]
        c.profile_enable(False)
        gbs = pd["bytes"] / 1e9 / (pd["ms"] / 1e3)
        getitem = {"decode_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "decode_ms": pd["ms"], "records": int(found.sum()),
                   "decoded_bytes": int(off[-1]), "lookup_ms": s.last_lookup_gpu_ms, "whole_call_ms": call_ms,
                   "whole_call_gbs": pd["bytes"] / 1e9 / (call_ms / 1e3),
                   "bytes_counted": "encoded read + decoded written; decode_* = decode kernels + arena memset (CUDA events), "
                                    "whole_call_* = pixiu_getitem_batch_dev wall time; `bench.py --mode getitem` is the full line"}
        # bit-exact round trip of a sample against the inputs
        from pixiu_b200.ctrl import split_doc
        hb = out[: int(off[-1])].cpu().numpy()
        keys, vals = synth.unpack(kd, ko), None
        vb = vd.tobytes()
        for i in list(range(0, n, max(n // 64, 1))):
            k, v = split_doc(hb[off[i]:off[i + 1]].tobytes())
            assert k == keys[i] and v == vb[vo[i]:vo[i + 1]], f"round trip mismatch at record {i}"
        getitem["roundtrip_sample_ok"] = True
    except Exception as e:  # the headline metric stands on its own
        getitem = {"error": repr(e)}

    # ---- CPU baseline: the reference on a bounded sample, every host core busy (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = args.ref_procs or (os.cpu_count() or 1)
        cpu = ref_all_cores(args.ref_pages, procs)
        one = _ref_worker((args.ref_pages, 2))          # one instance alone on the first pages of the workload itself
        cpu["single_instance"] = {"value": one[0] / one[1] / 1e6, "unit": UNIT, "cores": 1,
                                  "sample": f"first {args.ref_pages} pages of the workload ({one[0] / 1e6:.1f} MB raw)",
                                  "stored_over_raw_on_sample": one[2] / one[0]}
        cpu["host_cores"] = os.cpu_count()

    if rank == 0:
        total_raw = raw * world
        line = {
            "metric": METRIC, "value": total_raw * args.steps / (dev_ms / 1e3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C2: {args.pages} synthetic HTML-like pages x <=60KB (bytes 33..126), URL keys, batched setitem, per GPU",
                       "raw_bytes_per_gpu": raw, "records_per_gpu": n, "window_policy": args.window,
                       "windows_per_step": chunks_per_step, "stored_over_raw": ratio,
                       "l2": "inputs (~400 MB) and per-window scratch (~0.9 GB) exceed the 126 MB L2; no explicit flush",
                       "partitioning": "by key across ranks, no collective"},
            "wall_ms_per_step": dev_wall_ms / args.steps,
            "e2e": {"value": total_raw * args.steps / (e2e_ms / 1e3) / 1e6, "unit": UNIT,
                    "h2d_bytes_per_step": int(kd.nbytes + ko.nbytes + vd.nbytes + vo.nbytes), "d2h_bytes_per_step": int(8 * n),
                    "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "cpu_baseline": cpu, "getitem": getitem, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    c.free_prop()
    if dist is not None:
        dist.destroy_process_group()


def run_getitem(args, rank, world, local_rank):
    """BASELINE config[2] (C3): batched getitem decode of N records x 1 KB with deeply nested back references
    (or `--workload c2`: the HTML pages).  One step = lookup + decode of EVERY stored record of this rank.
      value : GB/s of (encoded bytes read + decoded bytes written), whole pixiu_getitem_batch_dev call timed with
              CUDA events on the store's stream (keys are host inputs, the decoded output stays in HBM)
      e2e   : pixiu_getitem_batch into a pinned HOST buffer (key H2D, decoded bytes D2H inside the timed region)
      roofline : the decode kernels alone (k_decode_tiles [+ k_resolve rounds]), algorithmic bytes / event time
    Multi-GPU: every rank holds its own records (partition by key), no collective."""
    import torch

    from pixiu_b200 import ctrl, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "c2":
        kd, ko, vd, vo = gen_corpus(args.pages, 2 + rank)
        wl = f"C2 decode: {args.pages} synthetic HTML-like pages x <=60KB per GPU, batched getitem of every record"
    else:
        kd, ko, vd, vo = synth.gen_nested(args.records, seed=3 + rank)
        wl = (f"C3: {args.records} records x 1 KB per GPU, record r = record r-1 with one swept byte mutated "
              f"(deeply nested back references) + self-periodic runs, batched getitem of every record")
    n = len(ko) - 1
    raw = int(ko[-1] + vo[-1])
    c = ctrl.PiXiuCtrl(device=local_rank, rotate_policy=ctrl.ROTATE_REFERENCE)
    rcs, _ = c.setitem_batch((kd, ko), (vd, vo))
    st = c.stats()
    ext = torch.cuda.ExternalStream(c.stream(), device=dev)
    cap = int(st.doc_bytes + 64)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    hout = torch.empty(cap, dtype=torch.uint8).pin_memory()
    hout_np = hout.numpy()
    alg = float(st.encoded_bytes + st.doc_bytes)

    def step_dev():
        return c.getitem_batch_dev((kd, ko), out.data_ptr(), cap)

    def step_host():
        return c.getitem_batch((kd, ko), out=hout_np)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        off, found = step_dev()
    assert found.all()
    # (a step is milliseconds of kernel time, not launch-bound: fast sampling; rank 0 only, NVML takes a driver-wide lock)
    sampler = ClockSampler(local_rank, first_delay=0.002, period=0.01) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = c.stats().kernel_launches
    dev_ms, dev_wall = timed(step_dev, args.steps)
    launches = c.stats().kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    step_host()
    e2e_ms, e2e_wall = timed(step_host, args.steps)
    # decode kernels alone
    c.profile_enable(True)
    step_dev()
    prof = c.profile()
    c.profile_enable(False)
    s = c.stats()
    pd = prof["decode"]
    peak, peak_kind = measured_peak()
    ach = pd["bytes"] / 1e9 / (pd["ms"] / 1e3)
    # DRAM traffic of k_decode_tiles over its algorithmic bytes from the committed `ncu --set full` captures
    # (profiles/r01_decode_tiles_c2_3000pages_ncu_raw.csv: 135.2 MB read + 77.9 MB written for 196.6 MB;
    #  profiles/r01_decode_tiles_c3_1Mrecords_ncu_raw.csv: 724.9 MB + 1,008.4 MB for 1,056.3 MB - deep chains re-read
    #  decoded sources that have left the L2), plus the arena memset (the decoded bytes written once more)
    ncu_ratio = {"c2": (135.17 + 77.93) / 196.57, "c3": (724.9 + 1008.4) / 1056.3}[args.workload]
    traffic = ncu_ratio * pd["bytes"] + float(st.doc_bytes)
    # bit-exact round trip of every record against the inputs (escape-free synthetic data)
    hb = out[: int(off[-1])].cpu().numpy()
    klen, vlen = np.diff(ko), np.diff(vo)
    ok = bool(np.array_equal(np.diff(off), klen + vlen + 4)
              and np.array_equal(synth.ragged_gather(hb, off[:-1], klen), kd)
              and np.array_equal(synth.ragged_gather(hb, off[:-1] + klen + 2, vlen), vd))
    if not ok:
        raise SystemExit("getitem bench: decoded records differ from the inputs")
    if rank == 0:
        line = {"metric": "getitem_decode_throughput", "value": alg * world * args.steps / (dev_ms / 1e3) / 1e9, "unit": "GB/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": wl, "records_per_gpu": n, "raw_bytes_per_gpu": raw, "decoded_bytes_per_gpu": int(st.doc_bytes),
                           "encoded_bytes_per_gpu": int(st.encoded_bytes), "chunks": int(st.chunks),
                           "bytes_counted": "encoded read + decoded written (SURVEY 8d)",
                           "l2": "decoded output (>= 1 GB) exceeds the 126 MB L2; no explicit flush"},
                "wall_ms_per_step": dev_wall / args.steps,
                "e2e": {"value": alg * world * args.steps / (e2e_ms / 1e3) / 1e9, "unit": "GB/s",
                        "h2d_bytes_per_step": int(kd.nbytes + ko.nbytes), "d2h_bytes_per_step": int(off[-1]) + 9 * n,
                        "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall / args.steps},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "k_decode_tiles(+k_resolve)", "achieved": ach, "peak": peak,
                             "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                             "algorithmic_bytes_per_launch": pd["bytes"], "avg_launch_ms": pd["ms"],
                             "launches": pd["launches"], "lookup_ms": s.last_lookup_gpu_ms},
                "roundtrip_all_records_ok": ok, "clocks": clocks, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    c.free_prop()
    if dist is not None:
        dist.destroy_process_group()


def run_lookup(args, rank, world, local_rank):
    """BASELINE config[3] (C4): batched contains over N URL keys with ~200-byte values (10 M keys = a ~2 GB compressed
    corpus); the query batch is N keys, 90 % present / 10 % absent, in random order.  One step = one batch.
      value : M keys/s, queries already packed in HBM (pixiu_contains_batch_dev), CUDA events on the store's stream
      e2e   : pixiu_contains_batch with HOST buffers (query H2D and found[] D2H inside the timed region)
      roofline : k_query_len/write + k_lookup, algorithmic bytes per query = q_len + depth x 7 + q_len"""
    import torch

    from pixiu_b200 import ctrl, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    n = args.keys
    kd, ko, vd, vo = synth.gen_urls_kv(n, seed=4 + rank, val_words=20)
    c = ctrl.PiXiuCtrl(device=local_rank, rotate_policy=ctrl.ROTATE_REFERENCE)
    t0 = time.perf_counter()
    rcs, _ = c.setitem_batch((kd, ko), (vd, vo))
    set_s = time.perf_counter() - t0
    st = c.stats()
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
    ext = torch.cuda.ExternalStream(c.stream(), device=dev)

    def step_dev():
        c.contains_batch_dev(d_qd.data_ptr(), d_qo.data_ptr(), n, d_found.data_ptr())

    def step_host():
        c._check(c._L.pixiu_contains_batch(c._h, n, p_qd.data_ptr(), p_qo.data_ptr(), found_h.ctypes.data_as(ctrl._u8p)))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(args.warmup):
        step_dev()
    if not np.array_equal(d_found.cpu().numpy().astype(bool), expect):
        raise SystemExit("lookup bench: found[] differs from the expected presence of the queries")
    sampler = ClockSampler(local_rank, first_delay=0.002, period=0.01) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = c.stats().kernel_launches
    dev_ms, dev_wall = timed(step_dev, args.steps)
    launches = c.stats().kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    step_host()
    if not np.array_equal(found_h.astype(bool), expect):
        raise SystemExit("lookup bench: host-path found[] differs")
    e2e_ms, e2e_wall = timed(step_host, args.steps)
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
    if rank == 0:
        line = {"metric": "contains_lookup_throughput", "value": n * world * args.steps / (dev_ms / 1e3) / 1e6, "unit": "Mkeys/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"C4: {n} URL keys x ~200 B values per GPU stored ({st.chunks} chunks, {st.encoded_bytes / 1e9:.2f} GB "
                                       f"compressed), batched contains of {n} keys, 90 % present / 10 % absent, random order",
                           "keys_per_gpu": n, "query_bytes": int(qd.nbytes), "mean_walk_depth": mean_depth,
                           "max_walk_depth_in_sample": int(depth.max()), "stored_over_raw": st.encoded_bytes / max(st.raw_bytes, 1),
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
                             "peak_kind": peak_kind, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                             "algorithmic_bytes_per_launch": alg, "avg_launch_ms": pl["ms"], "launches": pl["launches"],
                             "sector_granular_gbs": (pl["bytes"] + n * mean_depth * 32.0) / 1e9 / (pl["ms"] / 1e3),
                             "bytes_model": "2 x escaped query bytes + depth x 7 B per query (SURVEY 8d); sector-granular: depth x 32 B"},
                "found_matches_expected": True, "clocks": clocks, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    c.free_prop()
    if dist is not None:
        dist.destroy_process_group()


def run_sharded(args, rank, world, local_rank):
    """BASELINE config 5 style: ONE extended window sharded by record over the GPUs; every batch goes to
    all ranks, match lengths are MAX-reduced and leftmost candidates MIN-reduced by NCCL."""
    import torch
    import torch.distributed as dist

    from pixiu_b200 import ctrl, multigpu, shard, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kd, ko, vd, vo = gen_corpus(args.pages, 2)          # the same corpus on every rank
    n = len(ko) - 1
    raw = int(ko[-1] + vo[-1])
    bp = args.batch_pages
    batches = []
    for a in range(0, n, bp):
        idx = np.arange(a, min(n, a + bp))
        batches.append((shard.take_packed(kd, ko, idx), shard.take_packed(vd, vo, idx)))

    def one_pass():
        c = ctrl.PiXiuCtrl(device=local_rank, rotate_policy=ctrl.ROTATE_BYTES, window_bytes=args.window_bytes)
        c.mg_config(rank, world)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for kb, vb in batches:
            if world > 1:
                multigpu.setitem_sharded(c, kb, vb, device=dev)
            else:
                c.mg_setitem_begin(kb, vb)
                c.mg_setitem_mid()
                c.mg_setitem_end()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        st = c.stats()
        out = (dt, st.encoded_bytes / max(st.raw_bytes, 1), st.chunks, st.kernel_launches)
        c.free_prop()
        return out

    for _ in range(args.warmup):
        one_pass()
    res = [one_pass() for _ in range(args.steps)]
    t = torch.tensor([sum(r[0] for r in res)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t) / args.steps
        print(json.dumps({
            "metric": METRIC, "value": raw / sec / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C5 style: {args.pages} synthetic HTML-like pages, ONE window sharded over {world} GPU(s), "
                                   f"{args.window_bytes} window bytes per GPU, batches of {bp} pages replicated to all ranks",
                       "mode": "sharded window + NCCL all_reduce(MAX) of M / all_reduce(MIN) of leftmost candidates",
                       "raw_bytes": raw, "stored_over_raw": res[-1][1], "chunks": res[-1][2]},
            "gpu_launches": int(res[-1][3]), "collectives_per_step": 2 * len(batches) if world > 1 else 0}), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--ref-procs", type=int, default=0, help="reference instances run side by side (0 = one per host core)")
    ap.add_argument("--warmup-ref", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mode", default="partition", choices=["partition", "shard", "getitem", "lookup"],
                    help="partition: one store per GPU, corpus partitioned by key (default, weak scaling); "
                         "shard: one extended window sharded over the GPUs with NCCL reduces (config 5 style)")
    ap.add_argument("--batch-pages", type=int, default=256)
    ap.add_argument("--workload", default="c3", choices=["c2", "c3"], help="--mode getitem: which corpus to decode")
    ap.add_argument("--keys", type=int, default=10_000_000, help="--mode lookup: keys stored and queried per GPU")
    ap.add_argument("--records", type=int, default=1_000_000, help="--mode getitem --workload c3: records per GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.mode == "shard":
        run_sharded(args, rank, world, local_rank)
    elif args.mode == "getitem":
        run_getitem(args, rank, world, local_rank)
    elif args.mode == "lookup":
        run_lookup(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
