#!/usr/bin/env python
"""bench.py — MM-RCA fusion head fwd+bwd throughput (BASELINE.json metric, config 2).

  python bench.py --gpus N --steps K --warmup W            # B200 path (libmmrca.so kernels)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's algorithm on the host CPU cores

A "step" is one pass of the hot path over one batch of synthetic pooled features: forward, CrossEntropyLoss,
backward (run_one_epoch body restricted to the head, reference main_both.py:106-112), at batch 4096 per GPU
(weak scaling).  For N > 1 the script is launched once per rank by torchrun; every rank processes its own
batch and the flat head-gradient bucket is all-reduced once per step (plain data parallelism).

One JSON line on stdout (rank 0):
  value        samples/s with inputs already resident in HBM (device-timed, max over ranks)
  e2e          the same metric through the public host API with HOST buffers: pinned H2D of the features and
               labels (one staged record per batch, three slots deep on a copy stream; bf16 features when
               compute=bf16, MMRCA_FLAG_FEATURES_BF16) and D2H of loss + logits inside the timed region;
               e2e_fp32_features: the same with fp32 features
  gpu_eager_baseline   the reference's op sequence under PyTorch eager on the same GPU (fp32 and bf16 autocast)
  collective_check     N > 1: the peer-memory all-reduce against NCCL on the same bucket (untimed)
  roofline     dominant kernel: algorithmic FLOPs per launch / its CUDA-event duration, vs the measured
               bf16 tensor peak (MEASURED_PEAKS.json) — the roof SURVEY.md §8(d) assigns to the fused head
  cpu_baseline the oracle port of the reference timed on this box's host cores (N = 1 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_IMG, D_TXT, N_CLASSES = 1280, 768, 4
BATCH = 4096                       # BASELINE.json configs[1]
FLOPS_FWD = 2_895_872              # matmul FLOPs per sample, forward
FLOPS_FWD_BWD = 7_245_824          # forward + backward, features frozen (reference TL phase)
BYTES_PER_SAMPLE = 8_216           # features fp32 + label + logits (algorithmic HBM minimum)
L2_BYTES = 126 * 2 ** 20


def attn_flops(p_in, dkq, dv):
    return dict(proj=2 * 16 * p_in * (2 * dkq + dv), scores=2 * 16 * 16 * dkq, pv=2 * 16 * 16 * dv)


# algorithmic FLOPs per SAMPLE each kernel is responsible for; recompute is overhead
def kernel_flops():
    sa_i, sa_t, ca = attn_flops(80, 128, 96), attn_flops(48, 128, 96), attn_flops(96, 64, 48)
    f = {}
    f["attn_fwd<80,128,96,self>"] = sum(sa_i.values())
    f["attn_fwd<48,128,96,self>"] = sum(sa_t.values())
    f["attn_fwd<96,64,48,cross>"] = sum(ca.values())
    # backward kernels: attention backward (dP, dV = 2 pv; dQ, dK = 2 scores) + input grads where needed
    f["attn_bwd<80,128,96,self>"] = 2 * sa_i["pv"] + 2 * sa_i["scores"]
    f["attn_bwd<48,128,96,self>"] = 2 * sa_t["pv"] + 2 * sa_t["scores"]
    f["attn_bwd<96,64,48,cross>"] = 2 * ca["pv"] + 2 * ca["scores"] + ca["proj"]
    f["wgrad<80>"] = sa_i["proj"] / 3.0      # three launches (q, k, v) share the block's projection FLOPs
    f["wgrad<48>"] = sa_t["proj"] / 3.0
    f["wgrad<96>"] = ca["proj"] / 3.0
    f["classifier_fwd"] = 2 * 4 * 3584
    f["classifier_bwd"] = 2 * 2 * 4 * 3584
    f["cross_entropy"] = 0
    # bf16 tensor-core pipeline (mmrca_head_tc*.cuh).  Algorithmic = the reference's formulation (Q/K/V projections,
    # per-sample 16x16 attention), not the MMAs the kernels actually issue ("algorithmic work").
    bwd_sa = lambda a: a["proj"] + 2 * (a["scores"] + a["pv"])            # frozen features: no input gradient
    bwd_ca = lambda a: 2 * a["proj"] + 2 * (a["scores"] + a["pv"])
    f["sa_fwd_bf16"] = sum(sa_i.values()) + sum(sa_t.values()) + 2 * 4 * 2048
    f["ca_fwd_bf16"] = 2 * sum(ca.values()) + 2 * 4 * 1536
    f["ca_bwd_bf16"] = 2 * bwd_ca(ca) + 2 * 2 * 4 * 1536
    f["sa_bwd_bf16"] = bwd_sa(sa_i) + bwd_sa(sa_t)
    f["ce_feat"] = 2 * 4 * 2048                # dWf rows of the feature sources
    for k in ("prep_bf16", "finalize_bf16", "dropout_mask"):
        f[k] = 0
    return f


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_burst=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class NvmlClockSampler:
    """SM clock and throttle reasons polled through NVML from a thread while the timed region runs (the region
    is a few milliseconds long: nvidia-smi's 200 ms loop would not land a single sample inside it)."""
    HW_SLOWDOWN, SW_POWER_CAP, HW_THERMAL, SW_THERMAL = 0x8, 0x4, 0x40, 0x20

    def __init__(self, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        uuid = None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ent = vis.split(",")[index].strip()
            if ent.startswith("GPU-"):
                uuid = ent
            else:
                index = int(ent)
        self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(index)
        self.sm, self.reasons, self.run, self.thread = [], 0, False, None
        self.get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(self.get_reasons(self.h))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        self.run = True
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        self.run = False
        if self.thread:
            self.thread.join(timeout=1.0)
        names = [n for n, bit in (("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                                  ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP))
                 if self.reasons & bit]
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": names, "source": "nvml"}


def make_clock_sampler(index):
    try:
        return NvmlClockSampler(index)
    except Exception:
        return ClockSampler(index)


class ClockSampler:
    """Fallback: nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernel, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture (profiles/traffic.json, written by tools/ncu_summary.py traffic), if it was taken at this batch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        if int(d.get("batch", -1)) != batch:
            return None
        return d["kernels"].get(kernel)
    except (OSError, ValueError, KeyError):
        return None


def ncu_step_traffic(batch):
    """Sum of the per-launch DRAM bytes of every kernel of the step (profiles/traffic.json), if taken at this batch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        if int(d.get("batch", -1)) != batch:
            return None
        return float(sum(d["kernels"].values()))
    except (OSError, ValueError, KeyError):
        return None


def synth_batches(n, batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(n, batch, D_IMG, generator=g)
    txt = torch.randn(n, batch, D_TXT, generator=g)
    lab = torch.randint(0, N_CLASSES, (n, batch), generator=g)
    return img, txt, lab


# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own algorithm for this path on the host CPU: the oracle port (torch CPU restatement
    pinned to the reference by tests/golden; the Python reference itself cannot travel to the GPU box)."""
    if rank != 0:
        return
    import torch
    from oracle import mmrca_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = orc.init_head_params(seed=0)
    img, txt, lab = synth_batches(2, args.batch, 0)
    dm, ds = cpu_drop_mask(args.batch, args.dropout)
    step = lambda i: orc.head_loss_and_grads(p, img[i % 2], txt[i % 2], lab[i % 2], True, False, False,
                                             drop_mask=dm, drop_scale=ds)
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    val = args.steps * args.batch / dt
    emit(json.dumps({
        "impl": "reference", "metric": "mmrca_head_fwd_bwd_samples_per_s", "value": val, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MM_RCA --reverse fusion head fwd+CE+bwd, batch {args.batch}, features 1280+768, "
                               f"4 classes, train mode with dropout p={args.dropout} (fixed mask), backbones frozen "
                               "(BASELINE.json configs[1])"},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} steps of batch {args.batch} after {args.warmup} warm-up"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def cpu_drop_mask(batch, p):
    """A fixed keep mask for the CPU arms (drawing it is not part of the head; torch's own dropout would add the
    Philox draw to the CPU time, a fixed mask leaves it out in the CPU's favour)."""
    import torch
    if p <= 0:
        return None, 1.0
    g = torch.Generator().manual_seed(7)
    return (torch.rand(batch, 3584, generator=g) >= p).to(torch.uint8), 1.0 / (1.0 - p)


def cpu_baseline(batch, dropout, budget_s=12.0, max_steps=400):
    import torch
    from oracle import mmrca_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    p = orc.init_head_params(seed=0)
    img, txt, lab = synth_batches(1, batch, 0)
    dm, ds = cpu_drop_mask(batch, dropout)
    orc.head_loss_and_grads(p, img[0], txt[0], lab[0], True, False, False, drop_mask=dm, drop_scale=ds)
    n, t0 = 0, time.perf_counter()
    while n < max_steps and (time.perf_counter() - t0 < budget_s or n < 3):
        orc.head_loss_and_grads(p, img[0], txt[0], lab[0], True, False, False, drop_mask=dm, drop_scale=ds)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * batch / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} fwd+CE+bwd steps of batch {batch} ({dt:.1f} s) of the torch-CPU oracle port"}


def gpu_eager_baseline(batch, dropout, steps, warmup, dev):
    """The "existing Blackwell path" (SURVEY.md §2.2 / §8 d): the reference's own torch ops for this path - the oracle
    restatement, op for op what MM_RCA.forward + CrossEntropyLoss + backward() launch - run by PyTorch eager on the same
    B200, fp32 and under bf16 autocast, same batch / steps / warm-up, CUDA-event timed, inputs resident in HBM."""
    import torch
    from oracle import mmrca_oracle as orc
    p = {k: v.to(dev) for k, v in orc.init_head_params(seed=0).items()}
    img, txt, lab = synth_batches(2, batch, 0)
    img, txt, lab = img.to(dev), txt.to(dev), lab.to(dev)
    dm, ds = cpu_drop_mask(batch, dropout)
    dm = dm.to(dev) if dm is not None else None
    out = {}
    for name, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        def step(i):
            pp = {k: v.detach().requires_grad_(True) for k, v in p.items()}
            with torch.autocast("cuda", dtype=ctx, enabled=ctx is not None):
                logits = orc.head_forward(pp, img[i % 2], txt[i % 2], True, False, False, dm, ds)
            orc.cross_entropy(logits.float(), lab[i % 2]).backward()
        for i in range(max(warmup, 3)):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"value": batch / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms}
    out["what"] = ("PyTorch eager (torch %s) running the reference's op sequence for this path on the same GPU; fixed dropout "
                   "mask; %d steps after warm-up, CUDA events" % (torch.__version__, steps))
    return out


def collective_check(dp, step, dev, world):
    """N > 1, untimed: the same gradient bucket reduced through NCCL and through the one-shot peer-memory kernel
    (mmrca_peer_allreduce_mean): largest difference, and whether every rank ends with bit-identical bytes."""
    import torch
    import torch.distributed as dist
    if dp.peer is None:
        return {"skipped": "peer all-reduce not in use (%s)" % dp.collective}
    src = step.grads.flat.clone()
    src.add_(torch.arange(src.numel(), device=dev, dtype=torch.float32).mul_(1e-7 * (1 + dist.get_rank())))
    a, b = src.clone(), src.clone()
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    a.div_(world)
    dp.peer(b)
    torch.cuda.synchronize()
    err = (a - b).abs().max()
    ref = a.abs().max()
    bits = b.view(torch.int32).to(torch.int64)
    h = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=dev)).sum()])
    hmax, hmin = h.clone(), h.clone()
    dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    return {"max_abs_err": float(err), "max_abs_ref": float(ref), "bit_identical_across_ranks": bool((hmax == hmin).all()),
            "peer_status": int(dp.peer.status()),
            "bucket_floats": int(src.numel()), "against": "NCCL all_reduce(sum) / world"}


def bind_to_gpu_numa_node(local_rank):
    """N > 1: pin this rank (and, by first touch, its pinned host buffers) to the CPUs NVML reports as local to its GPU.
    The e2e measurement ships 16.8 MB per step and rank from host memory; eight ranks pulling through the wrong socket
    share one inter-socket link.  No-op when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ent = vis.split(",")[local_rank].strip()
            h = pynvml.nvmlDeviceGetHandleByUUID(ent) if ent.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(ent))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:      # noqa: BLE001
        pass
    return 0


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import garbage_classification_rca_b200 as g
    from garbage_classification_rca_b200 import _native as N
    from garbage_classification_rca_b200.training import HeadDataParallel

    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, args.warmup
    compute = N.COMPUTE_BF16 if args.compute == "bf16" else N.COMPUTE_FP32
    params = g.functional.init_head_parameters(dev, seed=0)
    co, fo = args.variant == "cross_only", args.variant == "features_only"
    if co or fo:
        params = g.functional.init_head_parameters(dev, seed=0, cross_attention_only=co, features_only=fo)
    step = g.HeadTrainStep(params, B, D_IMG, D_TXT, reverse=args.variant != "ca", cross_attention_only=co, features_only=fo,
                           compute=compute, drop_p=args.dropout)
    dp = HeadDataParallel(step, peer=not args.nccl)
    # inputs larger than L2: rotate over NB distinct batches
    per_batch = B * (D_IMG + D_TXT) * 4
    NB = max(2, -(-2 * L2_BYTES // per_batch))
    img_h, txt_h, lab_h = synth_batches(NB, B, 1234 + rank)
    img_d, txt_d, lab_d = img_h.to(dev), txt_h.to(dev), lab_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):      # zero_grad() is folded into the step's first kernel (MMRCA_FLAG_ZERO_GRADS)
        dp(img_d[i % NB], txt_d[i % NB], lab_d[i % NB], drop_seed=1000 + i, zero_grad=True)     # a fresh dropout mask every step

    # ---- value: device-resident inputs -------------------------------------------------------------
    for i in range(W):
        device_step(i)
    coll = collective_check(dp, step, dev, world) if world > 1 else None
    sampler = make_clock_sampler(local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
    barrier()
    N.kernel_launches(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        device_step(W + i)
    e1.record()
    barrier()
    launches = N.kernel_launches()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(step.loss.item())

    # ---- roofline: second pass with per-kernel CUDA events on the launch stream ----------------------
    per_kernel = {}
    N.timing_begin(launches + 64)
    for i in range(K):
        device_step(W + i)
    for name, t in N.timing_end(launches + 64):
        per_kernel.setdefault(name, []).append(t)
    torch.cuda.synchronize()

    # ---- e2e: host buffers, H2D on a copy stream three slots deep, D2H of loss + logits -----------------------
    # The bf16 pipeline's first act is to round the normalised features to bf16 (MMA operands), so with compute=bf16 the
    # host hands the features over AS bf16 (MMRCA_FLAG_FEATURES_BF16: what a backbone under bf16 autocast produces; half
    # the PCIe bytes).  The conversion of the synthetic host data is done once, outside the timed region: the timed
    # region starts from pinned host buffers of the dtype the call takes.  e2e_fp32_features: the same with fp32 features.
    NSLOT = 3
    out_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    out_logits = torch.empty(B, N_CLASSES, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)

    d2h_stream = torch.cuda.Stream(dev)

    def measure_e2e(fdt):
        # one pinned staging record per batch: [image features | text features | labels], shipped with ONE copy per step
        esz = torch.empty(0, dtype=fdt).element_size()
        off_t, off_l = B * D_IMG * esz, B * (D_IMG + D_TXT) * esz
        rec = off_l + B * 8
        host = torch.empty(NB, rec, dtype=torch.uint8).pin_memory()
        for j in range(NB):
            host[j, :off_t].copy_(img_h[j].to(fdt).reshape(-1).view(torch.uint8))
            host[j, off_t:off_l].copy_(txt_h[j].to(fdt).reshape(-1).view(torch.uint8))
            host[j, off_l:].copy_(lab_h[j].view(torch.uint8))
        raw = [torch.empty(rec, dtype=torch.uint8, device=dev) for _ in range(NSLOT)]
        slots = [(r[:off_t].view(fdt).view(B, D_IMG), r[off_t:off_l].view(fdt).view(B, D_TXT), r[off_l:].view(torch.int64))
                 for r in raw]
        ready = [torch.cuda.Event() for _ in range(NSLOT)]
        consumed = [torch.cuda.Event() for _ in range(NSLOT)]
        done = torch.cuda.Event()

        def h2d(i):
            sl = i % NSLOT
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[sl])
                raw[sl].copy_(host[i % NB], non_blocking=True)
                ready[sl].record(copy_stream)

        def e2e_loop(n):
            for sl in range(NSLOT):
                consumed[sl].record(main)
            for j in range(min(NSLOT - 1, n)):
                h2d(j)
            for i in range(n):
                if i + NSLOT - 1 < n:
                    h2d(i + NSLOT - 1)
                sl = i % NSLOT
                main.wait_event(ready[sl])
                main.wait_event(done)            # the previous step's loss / logits have left the device
                dp(*slots[sl], drop_seed=5000 + i, zero_grad=True)
                consumed[sl].record(main)
                with torch.cuda.stream(d2h_stream):      # results go back on their own stream, under the next step
                    d2h_stream.wait_event(consumed[sl])
                    out_loss.copy_(step.loss, non_blocking=True)
                    out_logits.copy_(step.logits, non_blocking=True)
                    done.record(d2h_stream)
            main.synchronize()
            d2h_stream.synchronize()

        e2e_loop(max(W, NSLOT))
        barrier()
        smp = make_clock_sampler(local_rank) if rank == 0 else None
        if rank == 0:
            smp.start()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(K)
        main.wait_event(done)
        t1.record()
        barrier()
        return t0.elapsed_time(t1), (smp.stop() if rank == 0 else None), rec

    ship_bf16 = compute == N.COMPUTE_BF16
    ms_e2e, clocks_e2e, h2d_bytes = measure_e2e(torch.bfloat16 if ship_bf16 else torch.float32)
    ms_e2e32, _, h2d_bytes32 = measure_e2e(torch.float32) if ship_bf16 else (ms_e2e, None, h2d_bytes)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e32], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e32 = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = world * B * K / (ms * 1e-3)
    e2e = world * B * K / (ms_e2e * 1e-3)
    kf = kernel_flops()
    share = {k: sum(v) / K for k, v in per_kernel.items()}           # ms per step per kernel name
    total_k = sum(share.values()) or 1.0
    dom = max(share, key=share.get)
    dom_ms = statistics.mean(per_kernel[dom])
    dom_flops = kf.get(dom, 0) * B
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    peak = peaks["bf16_sustained"]
    step_tflops = value / world * FLOPS_FWD_BWD / 1e12
    out = {
        "metric": "mmrca_head_fwd_bwd_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if compute == N.COMPUTE_FP32 else "bf16", "data": "synthetic",
        "config": {"workload": f"MM_RCA {'--reverse' if args.variant == 'rca' else 'plain cross-attention' if args.variant == 'ca' else '--reverse --cross_attention_only' if args.variant == 'cross_only' else '--features_only (fp32 kernels)'} fusion head fwd+CE+bwd, batch {B}/GPU, features 1280+768, "
                               f"4 classes, train mode with dropout p={args.dropout} (in-kernel seeded mask, new seed "
                               "every step), backbones frozen (BASELINE.json configs[1])",
                   "parallelism": f"dp{world}", "global_batch": world * B,
                   "collective": dp.collective if world > 1 else "none (1 GPU)",
                   "l2": f"inputs rotate over {NB} distinct batches ({NB * per_batch >> 20} MiB > 126 MiB L2)",
                   "compute": args.compute, "loss": loss_val,
                   "host_affinity": f"rank bound to the {numa_cpus} CPUs NVML reports local to its GPU" if numa_cpus else "default"},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4 + B * N_CLASSES * 4, "ms_per_step": ms_e2e / K,
                "features": "bf16 in pinned host memory (MMRCA_FLAG_FEATURES_BF16)" if ship_bf16 else "fp32 in pinned host memory",
                "copy_slots": NSLOT},
        "e2e_fp32_features": {"value": world * B * K / (ms_e2e32 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes32,
                              "d2h_bytes_per_step": 4 + B * N_CLASSES * 4, "ms_per_step": ms_e2e32 / K},
        "gpu_launches": launches,
        "clocks": dict(clocks, e2e_region=clocks_e2e),
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(dom, B), "peak_source": peaks["source"] + " bf16 sustained",
                     "kernel_ms": dom_ms, "kernel_share_of_step": share[dom] / total_k,
                     "timing": f"second pass of {K} steps with per-kernel CUDA events on the launch stream"},
        "roofline_step": {"bound": "tensor", "achieved": step_tflops, "peak": peak, "unit": "TFLOP/s",
                          "frac": step_tflops / peak, "flops_per_sample": FLOPS_FWD_BWD, "traffic": ncu_step_traffic(B),
                          "algorithmic_bytes": BYTES_PER_SAMPLE * B,
                          "hbm_view": {"achieved_gbs": value / world * BYTES_PER_SAMPLE / 1e9,
                                       "peak_gbs": peaks["hbm_gbs"]}},
        "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
    }
    # the streaming kernels of the step against the HBM roof: algorithmic bytes per sample (DESIGN.md kernel table)
    hbm_bytes = {"prep_feat": 8192 + 4608, "ce_feat": 4608}
    out["roofline_hbm_kernels"] = {
        k: {"achieved_gbs": hbm_bytes[k] * B / (share[k] * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
            "frac": hbm_bytes[k] * B / (share[k] * 1e-3) / 1e9 / peaks["hbm_gbs"], "bytes_per_sample": hbm_bytes[k]}
        for k in hbm_bytes if k in share}
    if coll is not None:
        out["collective_check"] = coll
    if world == 1 and compute == N.COMPUTE_BF16 and args.variant == "rca":
        # the fp32 SIMT kernels (the 1e-4 / per-tensor 1e-2 contract; feature gradients at fp32 accuracy) on the same
        # workload, a few steps: they are the other product path, and nobody should have to guess their speed
        step32 = g.HeadTrainStep(params, B, D_IMG, D_TXT, reverse=True, compute=N.COMPUTE_FP32, drop_p=args.dropout)
        for i in range(2):
            step32(img_d[i % NB], txt_d[i % NB], lab_d[i % NB], drop_seed=i, zero_grad=True)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(5):
            step32(img_d[i % NB], txt_d[i % NB], lab_d[i % NB], drop_seed=10 + i, zero_grad=True)
        f1.record()
        torch.cuda.synchronize()
        out["fp32_kernels"] = {"value": 5 * B / (f0.elapsed_time(f1) * 1e-3), "unit": "samples/s",
                               "ms_per_step": f0.elapsed_time(f1) / 5, "what": "MMRCA_COMPUTE_FP32 (SIMT) train step, same workload, 5 steps"}
    if world == 1 and not args.no_cpu_baseline:
        out["gpu_eager_baseline"] = gpu_eager_baseline(B, args.dropout, K, W, dev)
        out["cpu_baseline"] = cpu_baseline(B, args.dropout)
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------------------------------
def run_hier(args):
    """Secondary workload (not the BASELINE line): the hierarchical late-fusion head (--late_fusion=hierarchical,
    reference multimodal_model.py:729-818) fwd + CrossEntropyLoss + bwd at batch `--batch` on one B200, inputs resident
    in HBM (six pooled feature tensors, 32 KB / sample fp32), seeded dropout.  FLOPs: 2 * B * 8192 * 512 for the two
    hidden GEMMs forward, the same again for their weight gradients, + the 1024 -> 4 classifier."""
    import torch
    import garbage_classification_rca_b200 as g
    from garbage_classification_rca_b200 import _native as N
    from oracle import mmrca_oracle as orc      # parameter initialisation only (bench.py may use oracle/)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    p = orc.init_hier_params(seed=0)
    params = [p[n].to(dev) for n in g.functional.HIER_PARAM_NAMES]
    step = g.HierTrainStep(params, B, drop_p=args.dropout)
    gen = torch.Generator().manual_seed(5)
    NB = 2                                       # 2 x 134 MB of features > 126 MB L2
    feats = [[torch.randn(B, w, generator=gen).to(dev) for w in g.functional.HIER_SEGMENTS] for _ in range(NB)]
    labels = torch.randint(0, 4, (B,), generator=gen).to(dev)

    def one(i):
        step.zero_grad()
        step(feats[i % NB], labels, drop_seed=100 + i)

    for i in range(W):
        one(i)
    torch.cuda.synchronize()
    N.kernel_launches(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        one(W + i)
    e1.record()
    torch.cuda.synchronize()
    launches = N.kernel_launches()
    ms = e0.elapsed_time(e1)
    per_kernel = {}
    N.timing_begin(launches + 64)
    for i in range(K):
        one(W + i)
    for name, t in N.timing_end(launches + 64):
        per_kernel.setdefault(name, []).append(t)
    peaks = load_peaks()
    flops = {"hier_gemm": 2 * 8192 * 512 + 2 * 1024 * 4, "hier_wgrad": 2 * 8192 * 512, "hier_dh": 2 * 2 * 1024 * 4}
    share = {k: sum(v) / K for k, v in per_kernel.items()}
    dom = max(share, key=share.get)
    dom_ms = statistics.mean(per_kernel[dom])
    achieved = flops.get(dom, 0) * B / (dom_ms * 1e-3) / 1e12
    total = sum(flops.values())
    emit(json.dumps({
        "metric": "hierarchical_head_fwd_bwd_samples_per_s", "value": B * K / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1,
        "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Hierarchical fusion head fwd+CE+bwd, batch {B}, features 5888+2304, hidden 512+512, 4 classes, "
                               f"dropout p={args.dropout}, backbones frozen (SURVEY.md §8 a11 / f-1; secondary workload)",
                   "l2": "inputs rotate over 2 batches (268 MB > 126 MB L2)", "loss": float(step.loss.item())},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "kernel_ms": dom_ms, "traffic": None},
        "roofline_step": {"bound": "tensor", "achieved": B * K / (ms * 1e-3) * total / 1e12, "peak": peaks["bf16_sustained"],
                          "frac": B * K / (ms * 1e-3) * total / 1e12 / peaks["bf16_sustained"], "flops_per_sample": total},
        "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(share.items(), key=lambda kv: -kv[1])}}))

# ------------------------------------------------------------------------------------------------------
def run_full(args, rank, world, local_rank):
    """Secondary workload (BASELINE.json configs[2]): one full MM_RCA training step in the reference's transfer-learning
    phase — stock EfficientNetV2-M (480 x 480) + BERT-base (512 tokens) under bf16 autocast with frozen, random-init
    weights (no network for checkpoints), the B200 fusion head (bf16 pipeline) with CrossEntropyLoss, backward, one
    all-reduce of the head gradients, SGD step — at `--batch` samples per GPU (256 in BASELINE.json).  > 99.99 % of the
    FLOPs are the stock backbones; the line shows the drop-in module inside a real step."""
    import io
    from contextlib import redirect_stdout
    import torch
    import torch.distributed as dist
    from garbage_classification_rca_b200 import _native as N, multimodal_model as M
    from garbage_classification_rca_b200.training import CrossEntropyLoss, FusedSGD, PeerAllReduce, allreduce_mean_
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    torch.manual_seed(1 + rank)
    with redirect_stdout(io.StringIO()):
        m = M.MM_RCA(4, args.dropout, 0.0, 0.7, 256, "bert", B, True, False, False, pretrained=False, compute=N.COMPUTE_BF16)
    m = m.to(dev).train()
    # module path with ONE persistent gradient bucket (the backward kernels accumulate into it, p.grad are views), ONE
    # collective on it, ONE fused SGD launch over the flat parameter bucket (reference: torch.optim.SGD, main_both.py:548)
    fg = m.attach_flat_grads(flat_params=True)
    opt = FusedSGD(m._flat_params, fg, lr=1e-3, weight_decay=0.03)
    peer, collective = None, "none (1 GPU)"
    if world > 1:
        ok = torch.zeros(1, device=dev)
        try:
            peer = PeerAllReduce(fg.flat.numel(), dev)
            ok += 1
        except Exception:      # noqa: BLE001
            peer = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1:
            peer = None
        collective = "mmrca_peer_allreduce_mean (NVLink peer memory)" if peer is not None else "NCCL all_reduce"
    crit = CrossEntropyLoss()
    T = m.get_max_token_size()
    H, Wd = m.get_image_size()
    images = torch.randn(B, 3, H, Wd, device=dev)
    ids = torch.randint(0, 30522, (B, T), device=dev)
    mask = torch.ones_like(ids)
    labels = torch.randint(0, 4, (B,), device=dev)
    sample_ids = torch.arange(B)

    def one(cached=False):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = m(ids, mask, images, sample_ids=sample_ids if cached else None)
        loss = crit(logits.float(), labels)
        loss.backward()
        if world > 1:
            peer(fg.flat) if peer is not None else allreduce_mean_(fg.flat)
        opt.step()
        opt.zero_grad()
        return loss

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, out

    for _ in range(W):
        one()
    N.kernel_launches(reset=True)
    ms, loss = timed(one, K)
    launches = N.kernel_launches()
    # the head's share of that step: the same head work (hand-off output -> logits -> loss -> backward -> fused SGD) alone
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        m._images, m._input_ids, m._attention_mask = images, ids, mask
        feats = m.backbone_features()

    def head_only():
        l = crit(m.forward_features(*feats).float(), labels)
        l.backward()
        opt.step()
        opt.zero_grad()
        return l

    for _ in range(3):
        head_only()
    ms_head, _ = timed(head_only, 4 * K)
    # epochs >= 2 of the frozen phase with the feature cache (training.FeatureCache): the backbones are skipped
    m.enable_feature_cache(B)
    for _ in range(3):
        one(cached=True)
    ms_cached, _ = timed(lambda: one(cached=True), 4 * K)
    if rank == 0:
        emit(json.dumps({
            "metric": "mmrca_full_step_samples_per_s", "value": world * B * K / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"full MM_RCA --reverse step, EfficientNetV2-M {H}x{Wd} + BERT-base {T} tokens (stock torch, bf16 "
                                   f"autocast, frozen random-init), fused feature hand-off, B200 fusion head (bf16 pipeline, bf16 features), "
                                   f"flat gradient bucket + fused SGD, batch {B}/GPU (BASELINE.json configs[2]; secondary workload)",
                       "parallelism": f"dp{world}", "collective": collective, "loss": float(loss.item())},
            "head": {"ms_per_step": ms_head / (4 * K), "share_of_step": (ms_head / (4 * K)) / (ms / K),
                     "what": "logits + CrossEntropyLoss + backward + fused SGD on the hand-off's features, module path"},
            "cached_features": {"value": world * B * 4 * K / (ms_cached * 1e-3), "unit": "samples/s",
                                "ms_per_step": ms_cached / (4 * K),
                                "what": "the same step with training.FeatureCache (frozen phase, epochs >= 2: backbones skipped)"},
            "gpu_launches": launches}))
    if world > 1:
        dist.destroy_process_group()


def run_fusion(args, normalized):
    """Secondary workload (not the BASELINE line): the classic / normalized late-fusion heads (--late_fusion=classic |
    normalized, reference multimodal_model.py:489-579) fwd + CrossEntropyLoss + bwd at batch `--batch` on one B200, inputs
    resident in HBM, seeded dropout 0.6; --compute fp32 (1e-4 contract) or bf16 (Linear layers on the tensor cores)."""
    import torch
    from garbage_classification_rca_b200 import _native as N, functional as F
    from oracle import mmrca_oracle as orc      # parameter initialisation only (bench.py may use oracle/)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    p = orc.init_fusion_params(seed=0)
    step = F.FusionTrainStep([p[n].to(dev) for n in F.FUSION_PARAM_NAMES], B, normalized=normalized, drop_p=args.dropout,
                             compute=args.compute)
    gen = torch.Generator().manual_seed(3)
    NB = 8      # 8 x 32 MB of features: rotates over more than the L2
    img = [torch.randn(B, D_IMG, generator=gen).to(dev) for _ in range(NB)]
    txt = [torch.randn(B, D_TXT, generator=gen).to(dev) for _ in range(NB)]
    labels = torch.randint(0, N_CLASSES, (B,), generator=gen).to(dev)

    def one(i):
        step.zero_grad()
        step(img[i % NB], txt[i % NB], labels, drop_seed=100 + i)

    for i in range(W):
        one(i)
    torch.cuda.synchronize()
    N.kernel_launches(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        one(W + i)
    e1.record()
    torch.cuda.synchronize()
    launches = N.kernel_launches()
    ms = e0.elapsed_time(e1)
    per_kernel = {}
    N.timing_begin(launches + 64)
    for i in range(K):
        one(W + i)
    for name, t in N.timing_end(launches + 64):
        per_kernel.setdefault(name, []).append(t)
    H = 256
    flops = 3 * 2.0 * (D_IMG * H + D_TXT * H + 2 * H * H + H * N_CLASSES) - 2.0 * (D_IMG + D_TXT) * H   # no feature gradients
    peaks = load_peaks()
    emit(json.dumps({
        "metric": ("normalized" if normalized else "classic") + "_head_fwd_bwd_samples_per_s", "value": B * K / (ms * 1e-3),
        "unit": "samples/s", "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
        "dtype": args.compute, "data": "synthetic",
        "config": {"workload": f"{'Normalized' if normalized else 'Classic'} late-fusion head fwd+CE+bwd, batch {B}, features "
                               f"1280 + 768, hidden 256, 4 classes, dropout {args.dropout}; secondary workload",
                   "l2": f"inputs rotate over {NB} batches"},
        "gpu_launches": launches,
        "roofline_step": {"bound": "tensor" if args.compute == "bf16" else "fp32", "achieved": B * K / (ms * 1e-3) * flops / 1e12,
                          "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "flops_per_sample": flops,
                          "note": "launch-latency-bound at this size: ~30 launches of 4-20 us"},
        "kernels_ms_per_step": {k: round(sum(v) / K, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -sum(kv[1]))}}))


def run_token(args, rank, world, local_rank):
    """Secondary workload (BASELINE.json configs[4] / SURVEY.md §8 d cfg 5): the attention blocks on real token sequences,
    forward + backward (parameter gradients of the four blocks, input gradients of the cross blocks) + one fused SGD launch over
    the flat parameter bucket (the bf16 weight images are rebuilt from the updated fp32 parameters every step), `--batch` samples per GPU, data parallel with
    one all-reduce of the gradient bucket; `--token-forward-only`: inference replicas, no collective:
      SelfAttention(1024 -> 128 / 96) on ViT-L/16 tokens [B, 197, 1024], SelfAttention(768 -> 128 / 96) on RoBERTa tokens
      [B, 256, 768], and a ReverseCrossAttention(96 -> 64 / 48) on each sequence length (the reference asserts square
      attention, multimodal_model.py:93: the partner sequence is the block output of the neighbouring sample).
    The projection GEMM (K = 1024 / 768) is where this path meets the tensor roof: its achieved TFLOP/s is reported against
    the measured bf16 peak; FLOPs count the L valid tokens only (tiles are padded to 128 rows)."""
    import torch
    import torch.distributed as dist
    from garbage_classification_rca_b200 import _native as N, functional as F
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    g = torch.Generator().manual_seed(7 + rank)

    def block(d_q, d_kv, d_kq, d_v):
        def lin(o, i):
            k = 1.0 / i ** 0.5
            return [((torch.rand(o, i, generator=g) * 2 - 1) * k).to(dev), ((torch.rand(o, generator=g) * 2 - 1) * k).to(dev)]
        return lin(d_kq, d_q) + lin(d_kq, d_kv) + lin(d_v, d_kv) + [torch.ones(d_v, device=dev), torch.zeros(d_v, device=dev)]

    NB = 2      # 2 x (B x 197 x 1024 + B x 256 x 768) bf16: rotates over more than the L2 for B >= 128
    x_img = [torch.randn(B, 197, 1024, generator=g).bfloat16().to(dev) for _ in range(NB)]
    x_txt = [torch.randn(B, 256, 768, generator=g).bfloat16().to(dev) for _ in range(NB)]
    train = not args.token_forward_only
    blocks = [block(1024, 1024, 128, 96), block(768, 768, 128, 96), block(96, 96, 64, 48), block(96, 96, 64, 48)]
    # parameter bucket: one flat fp32 buffer (16-byte aligned tensors), the blocks read views of it; one fused SGD launch
    # updates all 32 tensors after the backward (mmrca_sgd_step), the gradient bucket below has the same layout
    offs, off = [], 0
    for blk in blocks:
        for t in blk:
            offs.append(off)
            off += (t.numel() + 3) // 4 * 4
    n_par = off
    flat_p = torch.zeros(n_par, device=dev)
    it_off = iter(offs)
    for blk in blocks:
        for j, t in enumerate(blk):
            o = next(it_off)
            v = flat_p[o:o + t.numel()].view_as(t)
            v.copy_(t)
            blk[j] = v
    sa_i = F.TokenAttention(blocks[0], B, 197, training=train, out_dtype=torch.bfloat16)
    sa_t = F.TokenAttention(blocks[1], B, 256, training=train, out_dtype=torch.bfloat16)
    # the self blocks write bf16 straight into samples 1 .. B of a [B + 1] buffer; sample 0 is a copy of sample B: samples
    # 1 .. B are the queries, samples 0 .. B - 1 the partner sequence (two contiguous views, no rolled copy)
    ext_i = torch.zeros(B + 1, 197, 96, dtype=torch.bfloat16, device=dev)
    ext_t = torch.zeros(B + 1, 256, 96, dtype=torch.bfloat16, device=dev)
    ca_i = F.TokenAttention(blocks[2], B, 197, reverse=True, training=train)
    ca_t = F.TokenAttention(blocks[3], B, 256, reverse=True, training=train)
    # gradient bucket: one flat fp32 buffer, the per-tensor views go to the backward calls; zeroed once per step
    flat_g = torch.zeros(n_par, device=dev)
    views, it_off = [], iter(offs)
    for blk in blocks:
        vs = []
        for t in blk:
            o = next(it_off)
            vs.append(flat_g[o:o + t.numel()].view_as(t))
        views.append(vs)
    d_ca_i = (torch.randn(B, 197, 48, generator=g) / (B * 197)).to(dev)
    d_ca_t = (torch.randn(B, 256, 48, generator=g) / (B * 256)).to(dev)

    # the image and the text branch are independent (self block -> cross block -> backward of both): each runs on its own
    # stream, so one branch's CTAs fill the SMs the other's last wave leaves idle (every attention kernel is ~3.5 waves
    # of one-CTA-per-SM work).  The per-kernel timing pass below runs them on one stream.
    s_img, s_txt = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def branch(sa, ca, x, ext, d_ca, v_sa, v_ca):
        if train:       # a training step follows an optimizer step: the bf16 weight images are rebuilt from the fp32 parameters
            sa.refresh_weights()
            ca.refresh_weights()
        sa(x, out=ext[1:])
        ext[0].copy_(ext[B])
        ca(ext[1:], ext[:B])
        if not train:
            return
        # backward: the cross block hands d(SA output) back (query side + the partner side), the self block stops at the
        # frozen backbone's tokens (no input gradient)
        dq, dkv = ca.backward(d_ca, v_ca, True, True)
        dq[:-1] += dkv[1:]      # d(SA output of sample b) = dq[b] + dkv[b + 1] (b is the partner of b + 1)
        dq[-1] += dkv[0]
        sa.backward(dq, v_sa)

    def compute(i, two_streams=True):
        if train:
            flat_g.zero_()
        jobs = ((s_img, (sa_i, ca_i, x_img[i % NB], ext_i, d_ca_i, views[0], views[2])),
                (s_txt, (sa_t, ca_t, x_txt[i % NB], ext_t, d_ca_t, views[1], views[3])))
        if two_streams:
            cur = torch.cuda.current_stream(dev)
            for st, job in jobs:
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    branch(*job)
            for st, _ in jobs:
                cur.wait_stream(st)
        else:
            for _, job in jobs:
                branch(*job)

    def finish():
        if train and world > 1:   # data parallel: one all-reduce of the flat gradient bucket (1.3 M floats) per step
            dist.all_reduce(flat_g)
        if train:                 # torch.optim.SGD(lr) on the whole bucket, one launch; the next step rebuilds the bf16 weight images
            N.check(N.lib().mmrca_sgd_step(flat_p.data_ptr(), flat_g.data_ptr(), None, n_par, 1e-3 / world, 0.0, 0.0, 0.0, 0, 0,
                                           torch.cuda.current_stream(dev).cuda_stream), "mmrca_sgd_step")

    graphs, graph_launches = None, 0

    def one(i, two_streams=True):
        if graphs is not None and two_streams:
            graphs[i % NB].replay()
        else:
            compute(i, two_streams)
        finish()

    for i in range(W):
        one(i)
    torch.cuda.synchronize()
    if not args.token_no_graph:
        # the step's ~50 launches on two streams (+ the library's side streams) captured once per input batch and replayed:
        # removes the host-side launch gaps (0.72 -> 0.69 ms); the collective and the optimizer launch stay outside
        cs = torch.cuda.Stream(dev)
        cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):
            compute(0)
        torch.cuda.current_stream(dev).wait_stream(cs)
        torch.cuda.synchronize()
        captured = []
        for j in range(NB):
            N.kernel_launches(reset=True)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                compute(j)
            graph_launches = N.kernel_launches()
            captured.append(gr)
        graphs = captured
        for i in range(W):
            one(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    N.kernel_launches(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        one(W + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = N.kernel_launches() + (graph_launches * K if graphs is not None else 0)
    ms = e0.elapsed_time(e1)
    per_kernel = {}
    N.timing_begin(launches + 64)
    for i in range(K):
        one(W + i, two_streams=False)
    recs = N.timing_end(launches + 64)
    torch.cuda.synchronize()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        # per step the projection launches come in a fixed order: SA image (N 352, K 1024), SA text (352, 768), then per
        # cross block Q (64, 96) and K|V (112, 96)
        proj = [r for r in recs if r[0] == "tok_proj"]
        shapes = [(197, 1024, 352), (256, 768, 352), (197, 96, 64), (197, 96, 112), (256, 96, 64), (256, 96, 112)]
        per_shape = {}
        for j, (name, t_ms) in enumerate(proj):
            per_shape.setdefault(shapes[j % 6], []).append(t_ms)
        for name, t_ms in recs:
            per_kernel.setdefault(name, []).append(t_ms)
        # weight-gradient GEMMs, host launch order per step: cross-image (q, k|v), self-image, cross-text (q, k|v), self-text
        wg = [t_ms for name, t_ms in recs if name == "tok_wgrad"]
        wg_big = [t for j, t in enumerate(wg) if j % 6 == 2] if train and len(wg) % 6 == 0 else []
        peaks = load_peaks()
        big = (197, 1024, 352)
        big_ms = statistics.mean(per_shape[big])
        big_flops = 2.0 * B * big[0] * big[1] * big[2]
        ach = big_flops / (big_ms * 1e-3) / 1e12
        ll = 197 * 197 + 256 * 256
        flops_sample = sum(2.0 * L * Kd * Nn for (L, Kd, Nn) in shapes) + 2.0 * ll * (128 + 96) + 2.0 * ll * (64 + 48)
        if train:
            # backward: weight gradients of every projection, input gradients of the cross blocks' projections, and per block
            # C, d(weights), dV (d_v wide) + dQ, dK (d_kq wide)
            flops_sample += sum(2.0 * L * Kd * Nn for (L, Kd, Nn) in shapes) + sum(2.0 * L * Kd * Nn for (L, Kd, Nn) in shapes[2:]) + \
                2.0 * ll * (3 * 96 + 2 * 128) + 2.0 * ll * (3 * 48 + 2 * 64)
        emit(json.dumps({
            "metric": "token_level_rca_blocks_fwd_bwd_samples_per_s" if train else "token_level_rca_blocks_fwd_samples_per_s", "value": world * B * K / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"token-level attention blocks {'forward + backward + fused SGD step' if train else 'forward'} (BASELINE.json configs[4]): SelfAttention on ViT-L/16 tokens "
                                   f"[{B},197,1024] and RoBERTa tokens [{B},256,768], ReverseCrossAttention 96->64/48 at L=197 and L=256; "
                                   f"batch {B}/GPU, " + ("data parallel: one NCCL all-reduce of the flat gradient bucket per step"
                                                         if train else "replicas (no collective)") + "; secondary workload",
                       "parallelism": f"dp{world}" if train else f"replicas x{world}", "l2": f"inputs rotate over {NB} batches",
                       "streams": "image and text branch on two CUDA streams" + ("" if graphs is None else ", captured into CUDA graphs (one per input batch) and replayed")},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "tok_proj [B*197 x 1024] x [1024 x 352]", "achieved": ach,
                         "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                         "kernel_ms": big_ms, "traffic": None,
                         "flops": "2 * B * 197 * 1024 * 352 (valid tokens; tiles are padded to 2 x 128 rows per sample)"},
            "roofline_step": {"bound": "tensor", "achieved": B * K / (ms * 1e-3) * flops_sample / 1e12,
                              "peak": peaks["bf16_sustained"], "frac": B * K / (ms * 1e-3) * flops_sample / 1e12 / peaks["bf16_sustained"],
                              "flops_per_sample": flops_sample},
            "roofline_wgrad": ({"bound": "tensor", "kernel": "tok_wgrad dW^T[1024 x 352] += X^T[1024 x B*197] G[B*197 x 352] (split-K slabs; "
                                                           "the reduce kernel is separate)",
                                "achieved": big_flops / (statistics.mean(wg_big) * 1e-3) / 1e12, "peak": peaks["bf16_sustained"],
                                "unit": "TFLOP/s", "frac": big_flops / (statistics.mean(wg_big) * 1e-3) / 1e12 / peaks["bf16_sustained"],
                                "kernel_ms": statistics.mean(wg_big)} if wg_big else None),
            "proj_ms_by_shape": {f"L{L}_K{Kd}_N{Nn}": round(statistics.mean(v), 4) for (L, Kd, Nn), v in per_shape.items()},
            "kernels_ms_per_step": {k: round(sum(v) / K, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -sum(kv[1]))}}))
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; fd 1 itself is pointed at stderr for the rest of the run
    so that native libraries (NCCL prints its version banner to stdout) cannot put other lines next to it."""
    if _JSON_FD is None:
        print(line, flush=True)
    else:
        os.write(_JSON_FD, (line + "\n").encode())


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--compute", default="bf16", choices=("fp32", "bf16"))
    ap.add_argument("--dropout", type=float, default=0.6, help="model_dropout (reference options.py:25 default 0.6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--token-forward-only", action="store_true", help="--workload token: forward only (inference)")
    ap.add_argument("--token-no-graph", action="store_true", help="--workload token: launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--nccl", action="store_true", help="N > 1: all-reduce the gradient bucket with NCCL instead of the "
                                                        "one-shot peer-memory kernel")
    ap.add_argument("--workload", default="mmrca", choices=("mmrca", "hierarchical", "classic", "normalized", "full", "token"),
                    help="mmrca: the BASELINE.json line (default); hierarchical: the second --late_fusion head (1 GPU); full: a whole "
                         "training step with the stock backbones (BASELINE.json configs[2], use --batch 256)")
    ap.add_argument("--variant", default="rca", choices=("rca", "ca", "cross_only", "features_only"),
                    help="late-fusion ablation of the mmrca workload (BASELINE.json configs[3]): --reverse (default), plain "
                         "cross-attention, --cross_attention_only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "hierarchical":
        if rank == 0:
            run_hier(args)
    elif args.workload in ("classic", "normalized"):
        if rank == 0:
            run_fusion(args, args.workload == "normalized")
    elif args.workload == "full":
        run_full(args, rank, world, local_rank)
    elif args.workload == "token":
        run_token(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
