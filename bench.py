#!/usr/bin/env python
"""Headline benchmark: images/s of the x4 SR forward (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 0..4]
                    [--model rfdn|imdn|rlfn|bsrn] [--batch B] [--size H W] [--dtype f16|f32]

--config N selects BASELINE.json configs[N] (default 1, the configuration the metric is quoted on):
  0  IMDN baseline, one 256x256 fp32 image (the reference's own CPU-runnable case; fp32 parity mode of the engine)
  1  RFDN baseline, batch 1, 256x256 fp16                                  <- default
  2  RFDN baseline, batch 32 of DIV2K-val-shaped LR images (synthetic; bucketed by shape, no padding), fp16
  3  RLFN (team04), 256x256 fp16, 8 images per GPU (batch 64 over 8 GPUs)
  4  BSRN (team18), 480x270 LR fp16, 16 images per GPU (batch 128 over 8 GPUs)
A step = one forward of one per-GPU batch.  Each rank runs the same per-rank workload on its own GPU
(independent images, no data-path collective: "weak" scaling); value = images of all ranks / max-over-
ranks device time.  One JSON line is printed by rank 0.  With N > 1 the run also checks and times the scatter /
gather path (ntire2022_esr_b200/sharded.py::forward_sharded over NCCL): `sharded_check` and `gathered`.

--impl reference times the reference's own CPU path (PyTorch ATen fp32, all host threads) through the
oracle port oracle/esr_oracle_torch.py on rank 0 (the pure-Python reference tree cannot travel to the GPU
box; the port calls the same ATen kernels its nn.Modules dispatch to).
"""
from __future__ import annotations

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

IDS = {"imdn": -1, "rfdn": 0, "rlfn": 4, "bsrn": 18, "fmen": 3}
WEIGHTS = {"imdn": "imdn_baseline", "rfdn": "rfdn_baseline", "rlfn": "team04_rlfn", "bsrn": "team18_bsrn", "fmen": "team03_fmen"}
RANGE = {"imdn": 1.0, "rfdn": 255.0, "rlfn": 255.0, "bsrn": 1.0, "fmen": 255.0}
L2_BYTES = 126 * 2 ** 20

# BASELINE.json configs[i] -> (model, per-GPU batch, (H, W) or "div2k", dtype, description)
CONFIGS = {
    0: ("imdn", 1, (256, 256), "f32", "IMDN baseline x4 SR forward, one 256x256 fp32 LR image (BASELINE.json configs[0])"),
    1: ("rfdn", 1, (256, 256), "f16", "RFDN baseline x4 SR forward, batch 1 per GPU, 256x256 LR -> 1024x1024, f16 storage (BASELINE.json configs[1])"),
    2: ("rfdn", 32, "div2k", "f16", "RFDN baseline x4 SR forward, batch 32 per GPU of DIV2K-val-shaped LR images (long side 510, synthetic, "
                                    "bucketed by shape, no padding), f16 storage (BASELINE.json configs[2])"),
    3: ("rlfn", 8, (256, 256), "f16", "RLFN (team04) x4 SR forward, 8 images per GPU (batch 64 over 8 GPUs), 256x256 LR, f16 storage (BASELINE.json configs[3])"),
    4: ("bsrn", 16, (270, 480), "f16", "BSRN (team18) x4 SR forward, 16 images per GPU (batch 128 over 8 GPUs), 480x270 LR, f16 storage (BASELINE.json configs[4])"),
}

# DRAM bytes per launch of the dominant kernel from committed `ncu --set full` captures (dram__bytes_read.sum +
# dram__bytes_write.sum), keyed by (model, batch, H, W, dtype, kernel); anything else reports null.
NCU_TRAFFIC = {
    ("rfdn", 1, 256, 256, "f16", "conv_tc"): (8.56e6, "profiles/r1_conv_tc_b1_ncu_summary.csv"),
    # fused RFDB chain (four 3x3 layers): dram__bytes_read 17.245 MB + dram__bytes_write 0.242 MB per launch, cold cache
    ("rfdn", 1, 256, 256, "f16", "conv_chain"): (17.49e6, "profiles/r2_conv_chain_b1_ncu_raw.csv"),
}


def div2k_shapes(n, seed=0):
    """DIV2K-val LR shapes (long side 510 = 2040 / 4), short side drawn from the validation set's
    distribution, 25 % portrait."""
    import random

    rnd = random.Random(seed)
    pool = [339] * 20 + [384] * 4 + [342] * 3 + [288] * 2 + [324, 408, 510]
    out = []
    for _ in range(n):
        short = rnd.choice(pool)
        out.append((510, short) if rnd.random() < 0.25 else (short, 510))
    return out


def load_weights(arch):
    import numpy as np

    with np.load(os.path.join(ROOT, "tests", "golden", "weights", WEIGHTS[arch] + ".npz")) as z:
        return {k: z[k] for k in z.files}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.idx], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 8]
        if not rows:
            rows = [r for _, r in self.rows if len(r) >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        sm = [float(r[1]) for r in rows]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def _cpu_forward(a, OT, w, xs):
    for x in xs:
        OT.forward(a.model, w, x)


def run_reference(a):
    """CPU arm: reference's PyTorch path via the torch-functional oracle port, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import esr_oracle_torch as OT

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    w = OT.prepare(load_weights(a.model))
    # a bounded sample of the per-GPU workload per step: images of it up to 4 x 256 x 256 LR pixels (the CPU does
    # 2-15 img/s at 256x256), at least one image
    full = make_inputs(a, torch.Generator().manual_seed(0), torch, torch.float32)
    xs, n_img, px = [], 0, 0
    for x in full:
        per = x.shape[2] * x.shape[3]
        take = min(x.shape[0], max(0, (4 * 256 * 256 - px) // per))
        if n_img == 0:
            take = max(take, 1)
        if take > 0:
            xs.append(x[:take].contiguous())
            n_img += take
            px += take * per
    for _ in range(min(a.warmup, 3)):
        _cpu_forward(a, OT, w, xs)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        _cpu_forward(a, OT, w, xs)
    dt = time.perf_counter() - t0
    val = n_img * a.steps / dt
    out = {"impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": a.gpus,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a),
           "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                            "sample": f"{a.steps} steps of {n_img} image(s) {[tuple(x.shape) for x in xs]} fp32 through "
                                      f"oracle/esr_oracle_torch.py (ATen CPU kernels, {cores} threads)"},
           "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(a, world=1):
    """Identical for both arms and every N (the driver compares it)."""
    h, wd = a.size
    desc = a.desc or f"{a.model.upper()} x4 SR forward, batch {a.batch} per GPU, {h}x{wd} LR, {a.dtype}"
    return {"workload": desc, "model_id": IDS[a.model], "batch_per_gpu": a.batch,
            "lr_size": "div2k-val shapes (long side 510)" if a.div2k else [h, wd], "weights": "reference model_zoo (pretrained)",
            "parallelism": "independent images per rank, no collective on the data path"}


def make_inputs(a, gen, torch, tdt):
    """One per-GPU batch: a list of (B_i, 3, H_i, W_i) tensors - a single entry except for the DIV2K-shaped config,
    where the 32 images are bucketed by shape (never padded: padding changes the result)."""
    if not a.div2k:
        h, wd = a.size
        return [(torch.rand(a.batch, 3, h, wd, generator=gen) * RANGE[a.model]).to(tdt)]
    buckets = {}
    for hw in div2k_shapes(a.batch):
        buckets[hw] = buckets.get(hw, 0) + 1
    return [(torch.rand(n, 3, h, wd, generator=gen) * RANGE[a.model]).to(tdt) for (h, wd), n in sorted(buckets.items())]


def cpu_baseline(a, budget_s=12.0):
    import torch
    from oracle import esr_oracle_torch as OT

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    w = OT.prepare(load_weights(a.model))
    x = make_inputs(a, torch.Generator().manual_seed(0), torch, torch.float32)[0][:1].contiguous()   # one image of the workload
    OT.forward(a.model, w, x)
    n, t0 = 0, time.perf_counter()
    while True:
        OT.forward(a.model, w, x)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 200:
            break
    return {"value": n / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{n} forwards of one {tuple(x.shape)} fp32 image of the workload (reference's ATen CPU path via "
                      f"oracle/esr_oracle_torch.py, {cores} threads, {dt:.1f} s)"}


def pin_to_gpu_numa_node(local):
    """Bind this rank's host threads (and with them its pinned staging buffers: first touch) to the NUMA node the
    GPU hangs off.  At N = 8 every rank moves ~20 GB/s of results to the host; buffers on the far socket halve that."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # no sysfs / nvml in this container: leave the affinity alone
        return {"error": type(e).__name__}


def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    from ntire2022_esr_b200 import Engine, _cabi, build_model
    from ntire2022_esr_b200.sharded import forward_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    tdt = torch.float16 if a.dtype == "f16" else torch.float32
    elt = 2 if a.dtype == "f16" else 4
    model = build_model(IDS[a.model], state_dict=load_weights(a.model)).eval().to(dev)
    eng = model.engine(dev)
    B = a.batch
    g = torch.Generator().manual_seed(1234 + rank)
    proto = make_inputs(a, g, torch, tdt)
    in_b = sum(x.numel() for x in proto) * elt
    out_b = in_b * 16
    # distinct input/output sets so that consecutive steps never find their data in L2
    nset = max(2, min(64, -(-2 * L2_BYTES // (in_b + out_b))))
    xs = [[x.to(dev) for x in (proto if k == 0 else make_inputs(a, g, torch, tdt))] for k in range(nset)]
    ys = [[torch.empty(x.shape[0], 3, 4 * x.shape[2], 4 * x.shape[3], dtype=tdt, device=dev) for x in xs[k]] for k in range(nset)]

    def step(i):
        for x, y in zip(xs[i % nset], ys[i % nset]):
            eng.forward(x, out=y)

    def barrier():
        if world > 1:
            dist.barrier()

    for i in range(max(a.warmup, 2 * nset)):  # every (input, output) pair gets its plan / CUDA graph built untimed
        step(i)
    torch.cuda.synchronize()
    t_w = time.perf_counter()           # >= 1 s of the same step so the SM clock has ramped before timing
    i = 0
    while time.perf_counter() - t_w < 1.0:
        for _ in range(8):
            step(i)
            i += 1
        torch.cuda.synchronize()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local])
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_c0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    # nvidia-smi needs ~1 s of load to produce samples: keep running the very same step for the clock record
    t_probe = time.perf_counter()
    i = 0
    while time.perf_counter() - t_probe < 1.2:
        for _ in range(8):
            step(i)
            i += 1
        torch.cuda.synchronize()
    # ---- the same K steps with three independent requests in flight (one engine handle, workspace and stream
    # each): what the engine's serving entry points do.  At batch 1 a forward is a chain of short kernels and
    # most SMs idle in the launch gaps / epilogue tails / small ESA kernels; another request's kernels fill them.
    n_pipe = 3
    engs = [eng] + [Engine(a.model, local).load_state_dict(load_weights(a.model)) for _ in range(n_pipe - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_pipe)]

    def pipe_steps(n):
        for i in range(n):
            k = i % n_pipe
            with torch.cuda.stream(streams[k]):
                for x, y in zip(xs[i % nset], ys[i % nset]):
                    engs[k].forward(x, out=y)

    pipe_steps(max(a.warmup, 6 * nset))     # plans / graphs of every (engine, input, output) triple
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for st in streams:
        st.wait_event(p0)
    pipe_steps(a.steps)
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        torch.cuda.current_stream().wait_event(ev)
    p1.record()
    torch.cuda.synchronize()
    barrier()
    pipe_ms = p0.elapsed_time(p1)
    t_c1 = time.perf_counter()
    del engs[1:]
    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region --------
    dtc = _cabi.DTYPE_F16 if a.dtype == "f16" else _cabi.DTYPE_F32
    nbuf = 6
    n_e2e = max(8, min(a.steps, 200))

    def e2e_run(submit, bufs_in, bufs_out):
        """n_e2e steps; step i reuses host buffer set i % nbuf, whose previous occupant (step i - nbuf) must be complete;
        the engine keeps 4 requests in flight."""
        for i in range(nbuf):
            last = [submit(xh, yh) for xh, yh in zip(bufs_in[i], bufs_out[i])]
        eng.host_wait(-1)
        barrier()
        t0 = time.perf_counter()
        tickets = []
        for i in range(n_e2e):
            if i >= nbuf:
                eng.host_wait(tickets[i - nbuf])
            for xh, yh in zip(bufs_in[i % nbuf], bufs_out[i % nbuf]):
                t = submit(xh, yh)
            tickets.append(t)
        eng.host_wait(-1)
        dt = time.perf_counter() - t0
        barrier()
        return dt

    hx = [[x.cpu().pin_memory() for x in make_inputs(a, g, torch, tdt)] for _ in range(nbuf)]
    hy = [[torch.empty(x.shape[0], 3, 4 * x.shape[2], 4 * x.shape[3], dtype=tdt).pin_memory() for x in hx[k]] for k in range(nbuf)]
    t_e2e = e2e_run(lambda xh, yh: eng.forward_host_async_ptr(xh.data_ptr(), yh.data_ptr(), xh.shape[0], xh.shape[2], xh.shape[3], dtc),
                    hx, hy)
    # the same with uint8 images in and out (the reference's run(): uint2tensor4 -> forward -> tensor2uint on the device):
    # a quarter of the fp16 output bytes cross PCIe
    t_u8 = None
    if a.dtype == "f16":
        ux = [[(torch.rand(x.shape[0], x.shape[2], x.shape[3], 3, generator=g) * 255).to(torch.uint8).pin_memory() for x in hx[k]] for k in range(nbuf)]
        uy = [[torch.empty(x.shape[0], 4 * x.shape[1], 4 * x.shape[2], 3, dtype=torch.uint8).pin_memory() for x in ux[k]] for k in range(nbuf)]
        t_u8 = e2e_run(lambda xh, yh: eng.forward_host_u8_async_ptr(xh.data_ptr(), yh.data_ptr(), xh.shape[0], xh.shape[1], xh.shape[2],
                                                                     RANGE[a.model], dtc), ux, uy)
    sampler.stop()
    # ---- N > 1: the scatter / gather path (a root-held batch over the GPUs of the box through NCCL) -------------
    shard = None
    if world > 1 and not a.div2k:
        h, wd = a.size
        n_all = world * B
        root_x = (torch.rand(n_all, 3, h, wd, generator=torch.Generator().manual_seed(99)) * RANGE[a.model]).to(tdt).to(dev) if rank == 0 else None
        y_all = forward_sharded(model, root_x, n_all, (3, h, wd), tdt, dev)
        check = "n/a"
        if rank == 0:
            # bit-equality with the same images run on rank 0 alone, image by image (what batch sharding promises)
            ok = all(torch.equal(model(root_x[k:k + 1])[0], y_all[k]) for k in range(0, n_all, max(1, n_all // 8)))
            check = "ok" if ok else "MISMATCH"
        n_g = max(4, min(a.steps // 4, 50))
        gout = torch.empty(n_all, 3, 4 * h, 4 * wd, dtype=tdt, device=dev) if rank == 0 else None
        for _ in range(2):
            forward_sharded(model, root_x, n_all, (3, h, wd), tdt, dev, out=gout)
        torch.cuda.synchronize()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(n_g):
            forward_sharded(model, root_x, n_all, (3, h, wd), tdt, dev, out=gout)
        g1.record()
        torch.cuda.synchronize()
        barrier()
        shard = (check, n_all, n_g, g0.elapsed_time(g1))
    t = torch.tensor([ms, t_e2e * 1e3, pipe_ms, (t_u8 or 0.0) * 1e3, shard[3] if shard else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, pipe_ms, u8_ms, gat_ms = (float(v) for v in t)
    if rank == 0:
        pk = peaks()
        # ---- roofline of the dominant kernel, timed live launch by launch (largest sub-batch of the step) ----
        prof = eng.profile_launches(xs[0][0], ys[0][0], reps=20)
        tot_ms = sum(p[2] for p in prof)
        by = {}
        for name, fl, m in prof:
            k = name.split(":")[0]
            d = by.setdefault(k, [0, 0.0, 0.0])
            d[0] += 1; d[1] += fl; d[2] += m
        dom = max(by, key=lambda k: by[k][2])
        n_l, fl, m = by[dom]
        ach = fl / (m * 1e-3) / 1e12 if m > 0 else 0.0
        x0 = xs[0][0]
        tr = NCU_TRAFFIC.get((a.model, x0.shape[0], x0.shape[2], x0.shape[3], a.dtype, dom))
        names = [n for x in xs[0] for n in eng.launch_names(x.shape[0], x.shape[2], x.shape[3], dtc)]
        launches = len(names)                                             # graph nodes per step (incl. the flag memset)
        kernels = sum(1 for n in names if not n.startswith("memset"))      # ... of which kernels of this repo
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_burst"], "frac_of_sustained_peak": ach / pk["tf_sust"],
                "traffic": tr[0] if tr else None,
                "traffic_source": tr[1] if tr else "no ncu --set full capture committed for this (workload, kernel)",
                "peak_source": pk["src"] + ", burst bf16 (the kernel's launches are timed one by one, in isolation)",
                "launches_per_step": n_l, "kernel_share_of_step": m / tot_ms if tot_ms else None,
                "whole_step_tflops": sum(p[1] for p in prof) / (tot_ms * 1e-3) / 1e12 if tot_ms else None,
                "algorithmic_gflop_per_step": sum(p[1] for p in prof) / 1e9,
                "per_kernel_ms": {k: round(v[2], 5) for k, v in by.items()}}
        cpu = cpu_baseline(a) if world == 1 else None
        val = world * B * a.steps / (ms * 1e-3)
        e2e_float = {"value": world * B * n_e2e / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": in_b,
                     "d2h_bytes_per_step": out_b, "steps": n_e2e,
                     "api": "esr_forward_host_async + esr_host_wait (C ABI, pinned host buffers, 4 requests in flight, each "
                            "on its own stream and workspace; every step copies its float input H2D and its float output D2H)"}
        # The reference's host-to-host pipeline is uint8 image -> uint2tensor4 -> forward -> tensor2uint -> uint8 image
        # (test_demo.py:420-435); that is the headline end-to-end figure for the fp16 engine.  The float-tensor variant
        # moves 2x the bytes back to the host and, at N = 8, runs into the box's host-memory write bandwidth
        # (profiles/r2_pcie_probe_n8.json: 125 GB/s for 8 concurrent D2H streams = 19.9 k img/s of 6.3 MB results).
        e2e_u8 = ({"value": world * B * n_e2e / (u8_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": in_b // elt,
                   "d2h_bytes_per_step": out_b // elt, "steps": n_e2e,
                   "api": "esr_forward_host_u8_async + esr_host_wait (C ABI, pinned host buffers, 4 requests in flight): uint8 HWC "
                          "image in, uint8 HWC image out - the reference's run() loop body (uint2tensor4 -> forward -> "
                          "tensor2uint, test_demo.py:420-435) with both conversions on the device; every step copies its "
                          "input H2D and its output D2H"} if u8_ms > 0 else None)
        e2e_main = e2e_u8 if e2e_u8 is not None else e2e_float
        out = {"metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": workload_config(a),
               "e2e": e2e_main, "e2e_float": e2e_float if e2e_main is not e2e_float else None,
               "pipelined": {"value": world * B * a.steps / (pipe_ms * 1e-3), "unit": "images/s", "requests_in_flight": n_pipe,
                             "note": "same K device-resident steps issued round-robin on 3 engine handles / streams; "
                                     "`value` above is the strict one-request-at-a-time number"},
               "gpu_launches": kernels * a.steps, "launches_per_step": launches, "kernels_per_step": kernels,
               "clocks": sampler.summary(t_c0, t_c1), "roofline": roof, "cpu_baseline": cpu,
               "l2_policy": f"{nset} distinct input/output sets rotated ({nset * (in_b + out_b) >> 20} MiB > 126 MiB L2); "
                            "engine workspace reused as in serving"}
        if shard:
            out["sharded_check"] = shard[0]
            out["gathered"] = {"value": shard[1] * shard[2] / (gat_ms * 1e-3), "unit": "images/s", "images_per_step": shard[1],
                               "steps": shard[2],
                               "note": "forward_sharded: a batch held by rank 0 is scattered over the ranks (NCCL send/recv), "
                                       "every rank runs its shard, the outputs are gathered on rank 0; device-timed, max over ranks"}
        if world > 1:
            out["host_affinity"] = numa
            out["note_reference_arm"] = "the reference arm times ONE CPU process on rank 0 for every N: value / reference at N > 1 compares N GPUs with one host"
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS), help="BASELINE.json configs[i] (default 1)")
    ap.add_argument("--model", default=None, choices=list(IDS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--size", type=int, nargs=2, default=None)
    ap.add_argument("--dtype", default=None, choices=["f16", "f32"])
    a = ap.parse_args()
    custom = any(v is not None for v in (a.model, a.batch, a.size, a.dtype))
    cfg = CONFIGS[a.config if a.config is not None else 1]
    a.div2k = cfg[2] == "div2k" and not custom
    a.desc = None if custom else cfg[4]
    a.model = a.model or cfg[0]
    a.batch = a.batch or cfg[1]
    a.size = list(a.size) if a.size else (list(cfg[2]) if cfg[2] != "div2k" else [339, 510])
    a.dtype = a.dtype or cfg[3]
    if a.steps is None:   # a few hundred milliseconds to a few seconds of device time whatever the workload
        a.steps = 400 if a.batch * a.size[0] * a.size[1] <= 4 * 256 * 256 else 60
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        if a.steps > 40:  # bounded CPU sample: the whole run must end within a few minutes
            a.steps = 40
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
