#!/usr/bin/env python
"""bench.py - pair registrations / second on synthetic [5000-keypoint, 32-d, 60-rotation] cloud pairs.

One "step" = one pass of the per-pair hot path (mutual matcher -> Des2R coarse rotations -> coarse-
rotation-guided RANSAC, 1000 iterations -> two weighted-Kabsch refinements) over a batch of B
independent synthetic pairs through the batched C-ABI engine (roreg_register_batch).  This is the
reference's default CLI configuration (`python Test.py`: --ET yohoc, mutual matcher) at BASELINE.json's
configs[1] size.  Nothing is skipped inside the timed region; the hypotheses are drawn and solved on
the device.

  value  - pairs/s with descriptors + keypoints already resident in HBM (B x 77 MB > L2, so every step
           streams its inputs from HBM: no L2 flush needed, stated in config.l2)
  e2e    - same metric through the reference-facing scene driver (roreg_b200.scene.register_scene = mutual.run +
           yohoc.run of the plugins for a whole dataset): a synthetic 60-cloud / 225-pair scene with the cloud
           reuse of the reference's test sets; the timed region reads the 60 cached descriptor files, copies
           them host->device, registers the pairs and WRITES the reference's per-pair files (match / scores /
           DR_index .npy, {id0}-{id1}.npz) and pre.log
  e2e_pair_upload / e2e_cloud_reuse4 (extras) - the batched C-ABI call fed from pinned host buffers: both clouds
           re-uploaded for every pair / every uploaded cloud used by 4 registrations
  roofline / cpu_baseline - see DESIGN.md "Measurement"

Multi-GPU (torchrun, one rank per GPU): pairs shard across ranks with no data-path collective
(weak scaling: B pairs per rank per step); the only collective is the final NCCL all_gather of the
[steps*B,4,4] poses, inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # --impl reference: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host thread it
    # can (round 1's N >= 2 reference lines ran on ONE core).  Must happen before NumPy / torch read the variables at import.
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=64)
    ap.add_argument("--n", type=int, default=5000, help="keypoints per cloud")
    ap.add_argument("--max-iter", type=int, default=1000)
    ap.add_argument("--nn-mode", type=int, default=int(os.environ.get("ROREG_NN_MODE", "4")),
                    help="0 = float32 difference form (reference arithmetic), 1-3 = tcgen05 3xTF32 Gram variants, 4 = one fp16 two-accumulator Gram per pair (default)")
    ap.add_argument("--score-mode", type=int, default=int(os.environ.get("ROREG_SCORE_MODE", "1")),
                    help="one-shot scoring arithmetic: 0 float64, 1 float32 pre-filter + exact float64 re-check (default; bit-identical overlaps, tests/test_gpu_parity.py::test_score_mode1_equals_float64_scoring)")
    ap.add_argument("--pipelined", type=int, default=int(os.environ.get("ROREG_PIPELINED", "1")),
                    help="1 (default): the resident-input loop uses roreg_register_batch_pipelined (RANSAC tail of batch i-1 beside the pooling of batch i; identical results, test_register_batch_pipelined_equals_serial); 0: one serial enqueue per batch")
    ap.add_argument("--value-only", type=int, default=0,
                    help="1 (kernel A/B experiments only): one untimed-quality step for the e2e / scene phases; their numbers are then meaningless")
    ap.add_argument("--corr-mode", type=int, default=int(os.environ.get("ROREG_CORR_MODE", "3")),
                    help="0 = FP32 CUDA-core Gram, 1-2 = tcgen05 3xTF32 Gram, 3 = fp16 two-accumulator Gram, operands from registers (default)")
    ap.add_argument("--scene-clouds", type=int, default=60, help="e2e scene: clouds (files read + uploaded once each)")
    ap.add_argument("--scene-pairs", type=int, default=225, help="e2e scene: pairs registered (3DMatch: 433 clouds / 1623 pairs)")
    ap.add_argument("--extras", type=int, default=1, help="1: also time the tensor-core workloads (yohoo estimator with the ET network; Match_ot matcher) as extra blocks")
    ap.add_argument("--extra-pairs", type=int, default=32, help="pairs per step of the yohoo extra workload")
    ap.add_argument("--net-passes", type=int, default=1, help="GEMM passes of the extra workloads' networks: 1 = TF32 (what the reference's cuDNN / cuBLAS path does on Ampere+), 3 = 3xTF32 (float32-class, the parity mode)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="bounded CPU-baseline sample (seconds of host work)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Multi-rank runs only: keep this rank's threads - hence the first touch of its pinned host buffers - on the NUMA node of its
    GPU (sysfs numa_node of the GPU's PCI device), so that 8 ranks do not pull 8 x 55 GB/s across the socket interconnect.
    Returns the previous affinity mask (restored before the CPU baseline) or None when anything is unavailable."""
    if os.environ.get("ROREG_BENCH_NUMA", "1") == "0":
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                       # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return old
    except Exception:
        return None


def make_inputs(B, n, rank):
    from roreg_b200 import synth
    prs = [synth.make_pair(1000 * (rank + 1) + p, n=n) for p in range(B)]
    desc = np.stack([x for pr in prs for x in (pr["feats0"], pr["feats1"])])
    keys = np.stack([x for pr in prs for x in (pr["keys0"], pr["keys1"])])
    pc = np.array([[2 * i, 2 * i + 1] for i in range(B)], np.int32)
    return prs, desc, keys, pc


# ------------------------------------------------------------------------------------------------
def cpu_pipeline(pr, tables, max_iter, ird, seed):
    """The oracle's restatement of the same pipeline on the host (reference arithmetic, NumPy)."""
    from oracle import roreg_oracle as O
    pps, sc = O.mutual_run(pr["feats0"], pr["feats1"])
    dr = O.rindex(pr["feats0"], pr["feats1"], pps, tables.perm)
    k0 = pr["keys0"][pps[:, 0]]; k1 = pr["keys1"][pps[:, 1]]
    T, recall, _ = O.yohoc_ransac(k0, k1, sc, dr, ird, max_iter, rng=np.random.RandomState(seed))
    return T


def time_c_port(prs, n_pairs, max_iter, ird, seconds):
    """Bounded sample of the OPTIMISED host port (oracle/oracle_c.c: C + pthreads hot loops, NumPy/LAPACK host logic; falls back
    to the NumPy oracle on one core when the C library is missing).  Not what a RoReg user runs - the stronger figure."""
    from roreg_b200 import group
    try:
        from oracle import oracle_c
        have_c = oracle_c.available()
    except Exception:
        have_c = False
    tables = group.load()
    t0 = time.perf_counter()
    done = 0
    while True:                                   # cycle over the pairs for ~`seconds` of host work
        i = done % n_pairs
        if have_c:
            oracle_c.register_pair(prs[i], tables, max_iter, ird, seed=done)
        else:
            cpu_pipeline(prs[i], tables, max_iter, ird, done)
        done += 1
        if time.perf_counter() - t0 >= seconds or done >= 64 * n_pairs:
            break
    dt = time.perf_counter() - t0
    if have_c:
        cores = oracle_c.threads()
        how = "oracle/oracle_c.c hot loops (C + pthreads, all host threads) + NumPy/LAPACK host logic"
    else:
        cores = 1
        how = "oracle/roreg_oracle.py (NumPy restatement; BLAS-free difference-form kernels run on one core)"
    return {"value": done / dt, "unit": "pairs/s", "cores": cores, "pairs": done, "seconds": dt,
            "sample": f"{done} pair registrations ({n_pairs} distinct pairs), {dt:.1f} s wall, {how}"}


def time_reference_ops(prs, n_pairs, max_iter, ird, seconds):
    """Bounded sample of the path with the REFERENCE'S OWN tensor operations on the host (oracle/torch_mirror.py: the chunked torch
    `pdist` + per-chunk min of utils/knn_search.py, the [K,32,60,60] gather + einsum of test/estimator.py:85-89, the NumPy yohoc
    RANSAC + refiner), all intra-op threads - what the reference executes when CUDA is absent ("run the reference's
    own classes with torch.set_num_threads(os.cpu_count())"; its real classes were probed at 0.12 pairs/s on 8 cores, this
    restatement runs 0.24 there)."""
    import torch
    from oracle import torch_mirror
    from roreg_b200 import group
    tables = group.load()
    t0 = time.perf_counter()
    done = 0
    while True:
        torch_mirror.register_pair(prs[done % n_pairs], tables.perm, max_iter, ird, done)
        done += 1
        if time.perf_counter() - t0 >= seconds or done >= 64 * n_pairs:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "pairs/s", "cores": int(torch.get_num_threads()), "pairs": done, "seconds": dt,
            "sample": f"{done} pair registration(s) ({n_pairs} distinct pairs), {dt:.1f} s wall, oracle/torch_mirror.py: the reference's "
                      "tensor operations on the host (chunked torch pdist, [K,32,60,60] gather + einsum, NumPy RANSAC), all intra-op threads"}


def cpu_baseline(prs, n_pairs, max_iter, ird, seconds=10.0):
    """`cpu_baseline` of the JSON line: value = the reference's own operations (time_reference_ops); the optimised C port is
    reported beside it as `optimised_c_port` (a stronger baseline than the reference itself: 6-7x faster)."""
    try:
        r = time_reference_ops(prs, n_pairs, max_iter, ird, seconds)
        out = {"value": r["value"], "unit": "pairs/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    except Exception as e:                         # torch CPU ops unavailable: fall back to the C / NumPy port as the value
        c = time_c_port(prs, n_pairs, max_iter, ird, seconds)
        return {"value": c["value"], "unit": "pairs/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
                "reference_ops_unavailable": repr(e)}
    try:
        c = time_c_port(prs, n_pairs, max_iter, ird, max(2.0, 0.5 * seconds))
        out["optimised_c_port"] = {"value": c["value"], "unit": "pairs/s", "cores": c["cores"], "sample": c["sample"]}
    except Exception as e:
        out["optimised_c_port"] = {"unavailable": repr(e)}
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU operations for the path (oracle/torch_mirror.py) on the host cores, rank 0 only."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))            # all host threads this process may run on
    prs, _, _, _ = make_inputs(max(1, args.cpu_sample_pairs), args.n, 0)
    per_step = max(1.0, min(args.cpu_seconds, 200.0 / max(1, args.steps + args.warmup)))   # whole arm within a few minutes
    timer, kind_note = time_reference_ops, None
    try:
        time_reference_ops(prs[:1], 1, min(args.max_iter, 50), 0.1, 0.0)                  # import + first-call costs outside the timed steps
    except Exception as e:
        timer, kind_note = time_c_port, repr(e)
    pairs = secs = 0.0
    last = None
    for s in range(args.warmup + args.steps):
        last = timer(prs, len(prs), args.max_iter, 0.1, per_step)
        if s >= args.warmup:
            pairs += last["pairs"]; secs += last["seconds"]
    val = pairs / secs
    cb = {"value": val, "unit": "pairs/s", "cores": last["cores"], "kind": "port",
          "sample": f"{args.steps} steps x ~{per_step:.1f} s bounded samples ({int(pairs)} pair registrations in {secs:.1f} s, cycling over "
                    f"{len(prs)} distinct pairs of the workload; per-pair rate, so the arm's pairs_per_step label equals the GPU arm's); the "
                    "reference's per-pair file I/O (3 re-reads of both 38 MB descriptor files, test/matcher.py:66-67, test/estimator.py:106-107) "
                    "is EXCLUDED - inputs are in memory; " + last["sample"]}
    if kind_note:
        cb["reference_ops_unavailable"] = kind_note
    else:
        try:                                                                              # once, outside the timed steps
            c = time_c_port(prs, len(prs), args.max_iter, 0.1, 4.0)
            cb["optimised_c_port"] = {"value": c["value"], "unit": "pairs/s", "cores": c["cores"], "sample": c["sample"]}
        except Exception as e:
            cb["optimised_c_port"] = {"unavailable": repr(e)}
    line = {"impl": "reference", "metric": "pair registrations/sec (5000 kpt, 60-rot)", "value": val, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": workload_config(args, args.pairs_per_step), "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, B):
    return {"workload": f"mutual matcher + Des2R + yohoc RANSAC ({args.max_iter} iters) + 2x refine, one pair = 2 clouds x "
                        f"{args.n} keypoints x [32,60] f32 descriptors (BASELINE configs[1], reference default CLI config)",
            "pairs_per_step": B, "keypoints": args.n, "max_iter": args.max_iter, "ransac_ird": 0.1,
            "parallelism": f"pairs sharded over {args.gpus} GPU(s), no data-path collective",
            "l2": "inputs larger than L2 (B x 77 MB streamed per step), no flush needed"}


def extra_workloads(args, ctx, peaks):
    """Tensor-core-heavy configurations of the same path, timed on rank 0 as EXTRA blocks of the JSON line (the headline stays the
    reference's default CLI pipeline):
      yohoo  - mutual -> Des2R -> ET network on the <= max_iter hypotheses that are scored (network/eqv_trans.py:119-138, only group
               element 0 of the head and its 13 + 60 inputs are computed) -> hypotheses -> one-shot RANSAC -> 2x refine, through
               pipeline.YohooEngine at --extra-pairs pairs per step (test/estimator.py:445-454);
      match_ot - the rotation-coherence matcher's forward (network/rot_coh_match.py:339-390) on one pair of 2 x n keypoints.
    Networks use random weights of the checkpoints' shapes (synth.random_weights); inputs are resident."""
    import torch
    from roreg_b200 import matchot, pipeline, synth
    out = {}
    n, H, Bx, npass = args.n, args.max_iter, args.extra_pairs, args.net_passes
    tf32_peak = 0.5 * peaks.get("bf16_tflops_sustained", 1400.0)
    prs = [synth.make_pair(5000 + i, n=n, with_fcgf=True, max_res_deg=2.0) for i in range(min(Bx, 4))]
    reps = -(-Bx // len(prs))                                  # the batch cycles over a few distinct pairs (host RAM / generation time)
    desc = ctx.dev(np.stack([x for pr in prs for x in (pr["feats0"], pr["feats1"])])); fcgf = ctx.dev(np.stack([x for pr in prs for x in (pr["fcgf0"], pr["fcgf1"])]))
    keys = ctx.dev(np.stack([x for pr in prs for x in (pr["keys0"], pr["keys1"])]), torch.float64)
    pc = ctx.dev(np.array([[2 * (i % len(prs)), 2 * (i % len(prs)) + 1] for i in range(Bx)], np.int32))
    try:
        eng = pipeline.YohooEngine(ctx, synth.random_weights("ET", 102), npass=npass, max_iter=H, ird=0.1, nn_mode=args.nn_mode)
        for w in range(2):
            eng.register(desc, fcgf, keys, pc, seed=w)
        torch.cuda.synchronize()
        steps = 5
        marks = []
        l0 = ctx.launches
        for s_ in range(steps):
            o = eng.register(desc, fcgf, keys, pc, seed=10 + s_, events=marks)
        torch.cuda.synchronize()
        per = len(marks) // steps
        stage = {}
        for s_ in range(steps):
            m = marks[s_ * per:(s_ + 1) * per]
            for (_, e0), (name, e1) in zip(m[:-1], m[1:]):
                stage[name] = stage.get(name, 0.0) + e0.elapsed_time(e1) / steps
        ms = marks[0][1].elapsed_time(marks[-1][1]) / steps
        nh = float(o["n_hyp"].double().mean().item())
        flops = Bx * nh * 2.0 * (13 * 128 * 256 * 60 + 13 * 256 * 512 * 13 + 13 * 512 * 256 + 256 * 512 + 512 * 128 + 128 * 4)
        gt = np.stack([prs[i % len(prs)]["gt"] for i in range(Bx)])
        err = float(np.abs(o["poses"].cpu().numpy()[:, :3] - gt).max())
        ach = flops / (stage["et_network"] * 1e-3) / 1e12
        out["workload_yohoo"] = {
            "value": Bx / (ms * 1e-3), "unit": "pairs/s", "pairs_per_step": Bx, "steps": steps, "ms_per_step": ms, "stage_ms_per_step": stage,
            "gpu_launches_per_step": (ctx.launches - l0) / steps, "hypotheses_per_pair": nh, "net_passes": npass, "max_abs_err_vs_gt": err,
            "roofline": {"kernel": "gemm_tc_kernel (ET network: implicit group-conv GEMMs + FC head)", "bound": "tensor", "achieved": ach,
                         "peak": tf32_peak * (1.0 if npass == 1 else 1.0 / 3.0), "unit": "TFLOP/s",
                         "frac": ach / (tf32_peak * (1.0 if npass == 1 else 1.0 / 3.0)), "traffic": None,
                         "note": "algorithmic flops = the pruned network (99.2 MFLOP per hypothesis: 60 group elements in layer 1, the 13 "
                                 "neighbours of g = 0 in layer 2, g = 0 alone afterwards; the reference computes 483.8 MFLOP) x hypotheses / "
                                 "CUDA-event time of the ET stage; peak = half the measured sustained dense bf16 rate (TF32), divided by the "
                                 "number of passes"},
            "note": "extra: yohoo estimator (ET network inside the timed region), inputs resident, random weights of the checkpoint shapes "
                    "(ET head biased to the identity quaternion); the batch cycles over " + str(len(prs)) + " distinct pairs"}
        del eng
    except Exception as e:                                   # an extra block must never take the headline down
        out["workload_yohoo"] = {"unavailable": repr(e)}
    try:
        pr = prs[0]
        f0 = ctx.dev(pr["feats0"]); f1 = ctx.dev(pr["feats1"])
        k0 = ctx.dev(pr["keys0"].astype(np.float32)); k1 = ctx.dev(pr["keys1"].astype(np.float32))
        mo = matchot.MatchOT(ctx, synth.random_weights("RM", 104), npass=npass)
        for w in range(2):
            mo.forward(f1, f0, k1, k0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches; steps = 3
        e0.record()
        for s_ in range(steps):
            m0, s0 = mo.forward(f1, f0, k1, k0)                  # NOTE THE SWAP: the network's source side is cloud id1 (test/matcher.py:192-197)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        n_launch = (ctx.launches - l0) / steps
        ms_graph = None
        try:                                                  # the same forward replayed from a CUDA graph (one cudaGraphLaunch per pair)
            mo.forward_graphed(f1, f0, k1, k0); torch.cuda.synchronize()
            e0.record()
            for s_ in range(steps):
                mg, sg = mo.forward_graphed(f1, f0, k1, k0)
            e1.record(); torch.cuda.synchronize()
            if bool((mg == m0).all().item()):
                ms_graph = e0.elapsed_time(e1) / steps
        except Exception:
            ms_graph = None
        best = min(ms, ms_graph) if ms_graph else ms
        out["workload_match_ot"] = {"value": 1e3 / best, "unit": "pairs/s", "ms_per_pair": best, "ms_per_pair_eager": ms, "ms_per_pair_cuda_graph": ms_graph,
                                    "keypoints": n, "steps": steps, "net_passes": npass,
                                    "gpu_launches_per_pair": n_launch, "matched": int((m0 >= 0).sum().item()),
                                    "note": "extra: Match_ot.forward (--RM matcher: 2 graph blocks, 4 R-indicators, 100 Sinkhorn iterations, mutual "
                                            "assignment) on one pair, inputs resident, random weights of the checkpoint shapes"}
    except Exception as e:
        out["workload_match_ot"] = {"unavailable": repr(e)}
    try:
        # all-pairs 60-rotation correlation (north_star kernel 1): max / argmax over the 60 rotations of the 1920-term correlation
        # of EVERY (n, m) keypoint pair of one cloud pair - one persistent tcgen05 kernel, the rotation loop inside
        from roreg_b200 import nets
        from roreg_b200.ops import _ptr, _stream
        pr = prs[0]
        res = {}
        for np_ in (1, 3):
            g = nets.GroupNets(ctx, np_)
            xh, xl = g.pack([ctx.dev(pr["feats1"])], [None], [0], None, n); yh, yl = g.pack([ctx.dev(pr["feats0"])], [None], [0], None, n)
            best = torch.empty((n, n), dtype=torch.float32, device=ctx.device); ba = torch.empty((n, n), dtype=torch.uint8, device=ctx.device)
            nn_ = torch.empty(n, dtype=torch.int32, device=ctx.device); nna = torch.empty(n, dtype=torch.int32, device=ctx.device)
            nd = torch.empty(n, dtype=torch.float32, device=ctx.device)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            for rep in range(3):
                e0.record()
                rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh), _ptr(xl), n, _ptr(yh), _ptr(yl), n, np_, _ptr(best), _ptr(ba), None, None, None, _stream())
                e1.record(); torch.cuda.synchronize()
                assert rc == 0
            ms = e0.elapsed_time(e1)
            rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh), _ptr(xl), n, _ptr(yh), _ptr(yl), n, np_, _ptr(best), _ptr(ba), _ptr(nn_), _ptr(nna), _ptr(nd), _stream())
            torch.cuda.synchronize()
            corr = pr["corr0"]; sel = np.where(corr >= 0)[0]
            inv = np.full(n, -1); inv[corr[sel]] = sel; have = inv >= 0
            res[np_] = (ms, float((nn_.cpu().numpy()[have] == inv[have]).mean()), float((nna.cpu().numpy()[have] == pr["a"]).mean()))
        flops = 2.0 * n * n * 1920 * 60
        peak1 = 0.5 * peaks.get("bf16_tflops", 1690.0)
        out["workload_allpairs"] = {
            "value": 1e3 / res[1][0], "unit": "pairs/s", "ms_per_pair_tf32": res[1][0], "ms_per_pair_3xtf32": res[3][0], "keypoints": n,
            "nn_accuracy_vs_planted": res[1][1], "rotation_accuracy_vs_planted": res[1][2], "gpu_launches_per_pair": 1,
            "roofline": {"kernel": "allpairs_tc_kernel<1>", "bound": "tensor", "achieved": flops / (res[1][0] * 1e-3) / 1e12, "peak": peak1, "unit": "TFLOP/s",
                         "frac": flops / (res[1][0] * 1e-3) / 1e12 / peak1, "traffic": None,
                         "note": "5.76 TFLOP (2 N M 1920 x 60 rotations) / CUDA-event time of the one launch; peak = half the measured BURST dense bf16 "
                                 "rate (TF32, kernel timed alone); ncu: tensor pipe 94.5 % active (profiles/r02_allpairs.txt)"},
            "note": "extra: 60-rotation correlation max / argmax for every keypoint pair of one cloud pair (a strict superset of what the reference "
                    "evaluates), one persistent kernel"}
    except Exception as e:
        out["workload_allpairs"] = {"unavailable": repr(e)}
    return out


def scene_e2e(args, ctx, rank, world, barrier):
    """The headline e2e: one whole synthetic scene per rank through scene.register_scene - descriptor files read from the cache
    directory, host->device copies, batched registration, the reference's per-pair files and pre.log written - timed from the call
    to its return (the background writer has finished by then).  Rep 0 warms up (pinned ring, workspace, page cache), rep 1 is timed."""
    import shutil
    import tempfile
    import types
    import torch
    from roreg_b200 import scene, synth
    from roreg_b200.test._common import CacheLayout
    n_clouds, n_pairs = (8, 12) if args.value_only else (args.scene_clouds, args.scene_pairs)
    cores = len(os.sched_getaffinity(0))
    readers = max(2, min(12, cores // max(1, world) - 1))        # file-reader threads of this rank (the ranks share the host's cores)
    readers = int(os.environ.get("ROREG_SCENE_READERS", readers))
    need = n_clouds * args.n * 7680 * 1.1
    root = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(cand).free > need + (1 << 30):
                root = tempfile.mkdtemp(prefix=f"roreg_scene_r{rank}_", dir=cand); break
        except Exception:
            pass
    if root is None:
        return {"unavailable": "no scratch directory with room for the scene's descriptor files"}
    try:
        sc = synth.SynthScene(7000 + rank, n_clouds=n_clouds, n_pairs=n_pairs, n=args.n, name="synth/bench_scene")
        sc.write_cache(root)
        cfg = types.SimpleNamespace(output_cache_fn=root, model_fn="", SO3_related_files=None, backbone="FCGF", bs_GF=1250, bs_ET=1000,
                                    RD=False, RM=False, match_n=0.5, ransac_ird=0.1, corr_mode=args.corr_mode)
        gt = np.stack([sc.get_transform64(a, b) for a, b in sc.pair_ids])
        dt = err = None
        for rep in range(2):
            np.random.seed(11)
            barrier()
            t0 = time.perf_counter()
            res = scene.register_scene(cfg, sc, keynum=args.n, max_iter=args.max_iter, batch_pairs=args.pairs_per_step, nn_mode=args.nn_mode,
                                       seed=rep, ctx=ctx, shard=False, readers=readers, writer_threads=2)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            err = float(np.abs(res["poses"][:, :3] - gt).max())
        lay = CacheLayout(cfg, sc, args.n)
        files = sum(len(fs) for _, _, fs in os.walk(lay.match_dir))
        out_bytes = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(lay.match_dir) for f in fs)
        kmean = float(res["n_matches"].mean())
        return {"seconds": dt, "pairs": len(sc.pair_ids), "clouds": n_clouds, "max_abs_err_vs_gt": err,
                "h2d_bytes": n_clouds * args.n * (7680 + 24),
                "d2h_bytes": len(sc.pair_ids) * (args.n * 8 + args.n * 4 + 4 + 128 + 4 + 8),
                "files_read": n_clouds, "files_written": files, "bytes_written": out_bytes, "matches_per_pair": kmean, "dir": os.path.dirname(root),
                "reader_threads": readers, "host_cores": cores}
    finally:
        shutil.rmtree(root, ignore_errors=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from roreg_b200 import ops
    torch.cuda.set_device(local)
    ctx = ops.Context(local)
    ctx.set_corr_mode(args.corr_mode)
    ctx.set_score_mode(args.score_mode)
    B, n, H = args.pairs_per_step, args.n, args.max_iter
    prs, desc_h, keys_h, pc_h = make_inputs(B, n, rank)
    old_affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    desc_pin = torch.from_numpy(desc_h).pin_memory(); keys_pin = torch.from_numpy(keys_h).pin_memory()
    pc = ctx.dev(pc_h)
    dev = ctx.device
    desc_d = [torch.empty_like(desc_pin, device=dev) for _ in range(2)]
    keys_d = [torch.empty_like(keys_pin, device=dev) for _ in range(2)]
    desc_d[0].copy_(desc_pin); keys_d[0].copy_(keys_pin)
    outs = [None, None]

    def step_resident(buf=0, seed=0):
        outs[buf] = ctx.register_batch(desc_d[buf], keys_d[buf], pc, max_iter=H, ird=0.1, seed=seed, nn_mode=args.nn_mode, out=outs[buf])
        return outs[buf]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- resident-input throughput ("value") ----------------
    # The timed region runs the library's default (serial) schedule; --pipelined 1 switches to the pipelined entry point.
    ctx.set_timing(False)
    for w in range(max(3, args.warmup)):
        step_resident(0, w)
    if args.pipelined:                                          # grows both workspace slots outside the timed region
        wouts = [None, None]
        for w in range(max(3, args.warmup)):
            wouts[w & 1] = ctx.register_batch_pipelined(desc_d[0], keys_d[0], pc, max_iter=H, ird=0.1, seed=w, nn_mode=args.nn_mode, out=wouts[w & 1])
        ctx.flush_batches()
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    poses_all = torch.empty((args.steps, B, 4, 4), dtype=torch.float64, device=dev)
    barrier()
    l0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if args.pipelined:
        # steady-state pipeline: batch s's poses are complete after call s+1 (or the flush); fill and drain are inside the timed region
        pouts = [None, None]
        for s in range(args.steps):
            pouts[s & 1] = ctx.register_batch_pipelined(desc_d[0], keys_d[0], pc, max_iter=H, ird=0.1, seed=100 + s, nn_mode=args.nn_mode,
                                                        out=pouts[s & 1])
            if s:
                poses_all[s - 1].copy_(pouts[(s - 1) & 1]["poses"])
        ctx.flush_batches()
        poses_all[args.steps - 1].copy_(pouts[(args.steps - 1) & 1]["poses"])
        o = pouts[(args.steps - 1) & 1]
    else:
        for s in range(args.steps):
            o = step_resident(0, 100 + s)
            poses_all[s].copy_(o["poses"])
    if world > 1:
        gathered = torch.empty((world,) + tuple(poses_all.shape), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(gathered, poses_all)       # the path's only collective: final gather of poses
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---------------- per-stage durations: a SERIAL replay of the same steps ----------------
    # With per-stage timing on, the library enqueues every stage on one stream and records a CUDA event after each (inside the
    # overlapped schedule a stage's event pair would not bracket one kernel alone).  Also gives the serial pairs/s.
    ctx.set_timing(True)
    stage_acc = {k: 0.0 for k in ctx.STAGES}
    stage_steps = max(8, min(args.steps, 24))
    step_resident(0, 7)
    torch.cuda.synchronize()
    es0 = torch.cuda.Event(enable_timing=True); es1 = torch.cuda.Event(enable_timing=True)
    es0.record()
    n_stage_samples = 0
    for s in range(stage_steps):
        o = step_resident(0, 100 + s)
        if s % 4 == 3 or s == stage_steps - 1:      # read the per-stage events (host wait on this step only)
            for k, v in ctx.stage_ms().items():
                stage_acc[k] += v
            n_stage_samples += 1
    es1.record()
    torch.cuda.synchronize()
    serial_value = B * stage_steps / (es0.elapsed_time(es1) * 1e-3)
    ctx.set_timing(False)

    # sanity: every registered pose must agree with the planted ground truth (a wrong-but-fast kernel is not a result)
    gt = np.stack([pr["gt"] for pr in prs])
    err = np.abs(poses_all[-1].cpu().numpy()[:, :3] - gt).max()
    ok = bool(err < 2e-2)

    # ---------------- end-to-end with host buffers ("e2e") ----------------
    copy_stream = torch.cuda.Stream(); comp = torch.cuda.current_stream()
    poses_pin = torch.empty((B, 4, 4), dtype=torch.float64).pin_memory()
    ready = [torch.cuda.Event(), torch.cuda.Event()]; done = [torch.cuda.Event(), torch.cuda.Event()]

    def h2d(buf):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[buf])                  # previous consumer of this buffer finished
            desc_d[buf].copy_(desc_pin, non_blocking=True); keys_d[buf].copy_(keys_pin, non_blocking=True)
            ready[buf].record(copy_stream)

    e2e_steps = 1 if args.value_only else max(4, min(args.steps, 12))
    e2e_reps = 1 if args.value_only else 2
    for b in (0, 1):
        done[b].record(comp)
    for rep in range(e2e_reps):                                        # rep 0 = warm-up, rep 1 = timed
        barrier()
        t0 = time.perf_counter()
        h2d(0)
        for s in range(e2e_steps):
            buf = s & 1
            if s + 1 < e2e_steps:
                h2d((s + 1) & 1)
            comp.wait_event(ready[buf])
            o = step_resident(buf, 500 + s)
            done[buf].record(comp)
            poses_pin.copy_(o["poses"], non_blocking=True)
        barrier()
        e2e_s = time.perf_counter() - t0
    # ---------------- end-to-end, scene access pattern (extra, not the headline e2e) ----------------
    # In the reference's real runs a cloud takes part in several pairs of its scene (3DMatch test set: 433 clouds, 1623 pairs ->
    # 7.5 pair-sides per cloud) and the plugin keeps a scene's clouds on the device (roreg_b200/test/_common.py: CloudCache).
    # Modelled here as: upload the 2B clouds of a step once, register every pair REUSE times (4 uses per cloud).
    REUSE = 4
    pc_scene = ctx.dev(np.tile(pc_h, (REUSE, 1)))
    out_scene = [None]
    poses_scene_pin = torch.empty((REUSE * B, 4, 4), dtype=torch.float64).pin_memory()

    def step_scene(buf, seed):
        out_scene[0] = ctx.register_batch(desc_d[buf], keys_d[buf], pc_scene, max_iter=H, ird=0.1, seed=seed, nn_mode=args.nn_mode, out=out_scene[0])
        return out_scene[0]

    scene_steps = 1 if args.value_only else max(3, min(args.steps, 6))
    for b in (0, 1):
        done[b].record(comp)
    for rep in range(e2e_reps):                                        # rep 0 = warm-up (also grows the workspace), rep 1 = timed
        barrier()
        t0 = time.perf_counter()
        h2d(0)
        for s in range(scene_steps):
            buf = s & 1
            if s + 1 < scene_steps:
                h2d((s + 1) & 1)
            comp.wait_event(ready[buf])
            o2 = step_scene(buf, 900 + s)
            done[buf].record(comp)
            poses_scene_pin.copy_(o2["poses"], non_blocking=True)
        barrier()
        scene_s = time.perf_counter() - t0
    ts = torch.tensor([scene_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    e2e_scene_val = world * REUSE * B * scene_steps / float(ts.item())
    scene_err = float(np.abs(o2["poses"][:B].cpu().numpy()[:, :3] - gt).max())
    sc_res = scene_e2e(args, ctx, rank, world, barrier)
    ctx.set_corr_mode(args.corr_mode)
    clocks = sampler.stop()          # sampled over the resident-input region and the host-buffer (e2e) regions
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * B * e2e_steps / float(te.item())
    h2d_bytes = desc_pin.numel() * 4 + keys_pin.numel() * 8
    d2h_bytes = poses_pin.numel() * 8
    if "seconds" in sc_res:
        tsc = torch.tensor([sc_res["seconds"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tsc, op=dist.ReduceOp.MAX)
        scene_val = world * sc_res["pairs"] / float(tsc.item())
    else:
        scene_val = None

    # ---------------- roofline of the dominant stage ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0); tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    stage_ms = {k: v / max(1, n_stage_samples) for k, v in stage_acc.items()}
    dom = max(stage_ms, key=stage_ms.get)
    kavg = float(o["n_matches"].double().mean().item())
    alg = {   # algorithmic bytes / flops per LAUNCH GROUP (= per step of B pairs), DESIGN.md "Measurement"
        "inv_pool": ("hbm", B * (2 * n * 7680 + 2 * n * 128)),
        "group_corr": ("hbm", B * kavg * (2 * 7680 + 12)),
        "nn": ("tensor", B * 2.0 * n * n * 32),
        "score_select": ("hbm", B * (kavg * 52 + H * 96)),
        "refine": ("hbm", B * kavg * 52 * 4),
        "hypotheses": ("hbm", B * (kavg * 4 + H * 96)),
        "compact": ("hbm", B * n * 16),
    }
    bound, amount = alg[dom]
    dur = stage_ms[dom] * 1e-3
    if bound == "hbm":
        ach = amount / dur / 1e9; peak = hbm_peak; unit = "GB/s"
    else:
        ach = amount / dur / 1e12; peak = tc_peak; unit = "TFLOP/s"
    traffic = None
    try:       # dram__bytes_read+write of the dominant stage's main kernel from the committed ncu --set full capture, scaled to B pairs
        tj = json.load(open(os.path.join(REPO, "profiles", "r02_traffic.json")))["kernels"]
        kname = {"nn": "nn_tc4_kernel", "group_corr": "group_corr_tc3_kernel",
                 "score_select": "ransac_score_pre_kernel" if args.score_mode == 1 else "ransac_score_kernel",
                 "inv_pool": "inv_pool_t4_kernel"}.get(dom)
        if kname in tj:
            traffic = tj[kname]["dram_bytes_per_launch"] / tj[kname]["pairs_per_launch"] * B
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "traffic": traffic, "peak_source": src, "stage_ms_per_step": stage_ms,
                "serial_schedule_pairs_per_s": serial_value,
                "fused_step": {"algorithmic_bytes": B * (2 * n * 7680 + 2 * n * 128 + kavg * (2 * 7680 + 12)), "ms": sum(stage_ms.values()),
                               "hbm_frac": B * (2 * n * 7680 + 2 * n * 128 + kavg * (2 * 7680 + 12)) / (sum(stage_ms.values()) * 1e-3) / 1e9 / hbm_peak,
                               "ms_overlapped": ms_max / args.steps,
                               "hbm_frac_overlapped": B * (2 * n * 7680 + 2 * n * 128 + kavg * (2 * 7680 + 12)) / (ms_max / args.steps * 1e-3) / 1e9 / hbm_peak,
                               "survey_8d_bytes": B * 77.6e6 * (n / 5000.0),
                               "hbm_frac_survey_8d": B * 77.6e6 * (n / 5000.0) / (sum(stage_ms.values()) * 1e-3) / 1e9 / hbm_peak,
                               "hbm_frac_survey_8d_timed": B * 77.6e6 * (n / 5000.0) / (ms_max / args.steps * 1e-3) / 1e9 / hbm_peak,
                               "note": "algorithmic_bytes = descriptors read by the pooling pass + the rows Des2R gathers for the K matches (two passes "
                                       "over HBM, DESIGN.md section 3); survey_8d_bytes = SURVEY.md 8(d)'s 77.6 MB per pair (every descriptor byte "
                                       "counted once); ms = serial sum of the per-stage events, ms_overlapped / *_timed = the timed region's own "
                                       "time per step (pipelined schedule: the tail of batch i-1 beside the pooling of batch i)"},
                "note": "algorithmic bytes/flops per step (B pairs) / CUDA-event duration of that stage, events recorded by the library on the "
                        "launch stream in a replay of the timed steps with per-stage timing on (same serial schedule as the timed region unless "
                        "--pipelined 1; fused_step.ms_overlapped = the timed region's own ms per step, events included in neither)"}

    if old_affinity is not None:
        try:
            os.sched_setaffinity(0, old_affinity)
        except Exception:
            pass
    if rank == 0:
        cb = cpu_baseline(prs, min(args.cpu_sample_pairs, B), H, 0.1, args.cpu_seconds) if (world == 1 and args.cpu_sample_pairs > 0) else None
        line = {"metric": "pair registrations/sec (5000 kpt, 60-rot)", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (matcher, correlation) + f64 (RANSAC, Kabsch)",
                "data": "synthetic", "config": workload_config(args, B), "clocks": clocks,
                "e2e": ({"value": scene_val, "unit": "pairs/s", "h2d_bytes_per_step": sc_res["h2d_bytes"], "d2h_bytes_per_step": sc_res["d2h_bytes"],
                         "steps": 1, "seconds": sc_res["seconds"], "pairs_per_step": sc_res["pairs"], "clouds": sc_res["clouds"],
                         "files_read": sc_res["files_read"], "files_written": sc_res["files_written"], "bytes_written": sc_res["bytes_written"],
                         "matches_per_pair": sc_res["matches_per_pair"], "max_abs_err_vs_gt": sc_res["max_abs_err_vs_gt"], "scratch": sc_res["dir"],
                         "reader_threads_per_rank": sc_res["reader_threads"], "host_cores": sc_res["host_cores"],
                         "h2d_gb_per_s_per_rank": sc_res["h2d_bytes"] / sc_res["seconds"] / 1e9,
                         "note": "through roreg_b200.scene.register_scene (the plugins' mutual.run + yohoc.run for a whole dataset), one scene per "
                                 "rank: timed region = read the cached descriptor files (one per cloud) -> pinned ring -> HBM, register every pair "
                                 "in batches as its clouds arrive, device->host results, write match / scores / DR_index .npy + .npz per pair "
                                 "and pre.log; a step = one scene; warm-up = one untimed pass over the same scene.  Host-bound (file reads "
                                 "are memcpy out of the page cache / tmpfs, the ranks share the host's cores): N-GPU values scale with the "
                                 "host, not with the GPUs"}
                        if scene_val is not None else
                        {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                         "note": "scene e2e unavailable (" + sc_res.get("unavailable", "?") + "): per-pair upload figure instead"}),
                "e2e_pair_upload": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                                    "steps": e2e_steps, "h2d_gb_per_s_per_rank": h2d_bytes * e2e_steps / e2e_s / 1e9,
                                    "note": "extra (round 1's headline): roreg_register_batch fed from pinned host buffers, BOTH clouds of every "
                                            "pair re-uploaded each step (no real run does that: 433 clouds / 1623 pairs on 3DMatch); PCIe-bound"},
                "e2e_cloud_reuse4": {"value": e2e_scene_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes * REUSE,
                                     "pairs_per_step": REUSE * B, "steps": scene_steps, "max_abs_err_vs_gt": scene_err,
                                     "note": f"extra: same pinned buffers, every uploaded cloud used by {REUSE} registrations per step, no files"},
                "gpu_launches": int(launches), "roofline": roofline, "pose_check": {"max_abs_err_vs_gt": float(err), "ok": ok},
                "nn_mode": args.nn_mode, "corr_mode": args.corr_mode, "score_mode": args.score_mode, "pipelined": args.pipelined}
        if args.value_only:
            line["value_only"] = True; line["e2e"]["note"] = "--value-only run: the e2e figures were not measured properly (tiny scene, one step)"
        if cb:
            line["cpu_baseline"] = cb
        if args.extras and not args.value_only and world == 1:
            line.update(extra_workloads(args, ctx, peaks))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
