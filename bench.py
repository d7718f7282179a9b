#!/usr/bin/env python
"""bench.py — R(2+1)D-34 32x112x112 clips/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path (oracle/)

Workload at every N: BASELINE.json configs[1] — R(2+1)D-34 inference, 32x112x112 clips, 101 classes, batch 48 per
GPU, bf16 activations / fp32 accumulate, synthetic U[0,1) clips, Xavier random-init weights.  N > 1 shards by batch
(weak scaling, one process per GPU, no data-path collective: inference replicas).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_CLIP_FWD = 304.711020544      # 2*M*N*K over unpadded conv dims, R34 32x112^2 (oracle.conv_flops)
GFLOP_PER_CLIP_TRAIN = 912.8            # fwd + dgrad + wgrad, minus the stem dgrad
TRAIN_BATCH_PER_GPU = 4                 # BASELINE configs[2]
# dram__bytes_read.sum + dram__bytes_write.sum of one unit2p1_fused_kernel launch at batch 48 (ncu --set full); None until captured
FUSED_UNIT_DRAM_BYTES_B48 = 1.521e9   # profiles/r01z_ncu_unit2p1_fused_is.txt: 923 MB read + 598 MB written (algorithmic 617 + 617 MB: halo rows re-read)
MODEL_DEPTH, NUM_CLASS, T, HW = 34, 101, 32, 112
BATCH_PER_GPU = 48
C4_BATCH_PER_GPU, C4_T, C4_NUM_CLASS = 16, 16, 63     # BASELINE configs[3] (Meitu shape, train_simple_r3d.py:336,341)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d["bf16_tflops_sustained"], burst=d["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, burst=1590.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def synthetic_clips(n, seed=123):
    import numpy as np
    return np.random.default_rng(seed).random((n, 3, T, HW, HW), dtype=np.float32)


def synthetic_clips_u8(n, seed=123, t=None):
    """The same U[0,1) clips as decoded 8-bit frames (N, T, H, W, 3): what a video decoder hands over (1 byte per value)."""
    import numpy as np
    x = np.random.default_rng(seed).random((n, 3, t or T, HW, HW), dtype=np.float32)
    return np.ascontiguousarray((x * 255.0).astype(np.uint8).transpose(0, 2, 3, 4, 1))


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # data/ucf101.py:124-128


def oracle_params():
    from oracle import r2plus1d as orc
    return orc.randomize_bn(orc.init_params(MODEL_DEPTH, NUM_CLASS, seed=0), seed=1)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(clips_per_step, steps, warmup):
    """Times the CPU restatement (oracle.Net, torch-CPU/oneDNN, all host threads) of the same forward pass.
    torch.distributed.run exports OMP_NUM_THREADS=1 to its children: the thread count is set explicitly so that the
    reference arm uses every host core at every N."""
    import torch
    from oracle import r2plus1d as orc
    torch.set_num_threads(host_cores())
    net = orc.Net(oracle_params(), MODEL_DEPTH, (T // 8, HW // 16, HW // 16))
    x = synthetic_clips(clips_per_step)
    with torch.no_grad():
        for _ in range(warmup):
            net.forward(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            net.forward(x)
        dt = time.perf_counter() - t0
    return clips_per_step * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    clips = 2
    rate, sec_per_step, threads = cpu_reference_rate(clips, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "r2plus1d34_32x112_inference_clips_per_s", "value": rate, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "R(2+1)D-34 inference 32x112x112, 101 classes (BASELINE configs[1])",
                   "clips_per_step": clips, "note": "MXNet is not installable here; this is the oracle's torch-CPU "
                   "restatement of model/R2Plus1.py, fp32, on the host cores"},
        "cpu_baseline": {"value": rate, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": "%d clips/step x %d steps, R34 32x112x112 fp32" % (clips, args.steps)},
        "e2e": {"value": rate, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    from fastvideotagging_b200 import build
    build.build()
    from fastvideotagging_b200.model import R2Plus2D

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")          # keep stdout to the one JSON line (no NCCL version banner)
    if world > 1:
        # NCCL kernels on a high-priority stream: a gradient bucket that becomes ready is reduced while the remaining
        # data / weight gradients keep the SMs busy, instead of queueing behind them
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=dev, pg_options=opts)
        except (AttributeError, TypeError):
            dist.init_process_group("nccl", device_id=dev)

    def note(msg):
        if os.environ.get("FVT_BENCH_VERBOSE"):
            sys.stderr.write("[bench rank %d] %s\n" % (rank, msg))
            sys.stderr.flush()

    peaks = load_peaks()
    batch = args.batch
    net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
    net.load_param_dict(oracle_params())
    net.eval()
    net.set_input_normalization(IMAGENET_MEAN, IMAGENET_STD, scale=1.0 / 255.0)     # used by uint8 clip batches only

    if args.input == "u8":
        # decoded 8-bit frames; crop/normalise/unfold run fused into the stem's input transform (fvt_clip_unfold_u8)
        x_host = torch.from_numpy(synthetic_clips_u8(batch, seed=123 + rank)).pin_memory()
    else:
        x_host = torch.from_numpy(synthetic_clips(batch, seed=123 + rank)).pin_memory()
    x_dev = x_host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        plan = net._inference_plan(x_dev)
        n_launches = plan.launches
        # ---------------- device-resident throughput ("value")
        for _ in range(max(args.warmup, 3)):
            logits = net(x_dev)
        barrier()
        note("inference warm-up done")
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            logits = net(x_dev)
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)

        # ---------------- end-to-end through the public API with host buffers ("e2e")
        # double-buffered: the H2D copy of step i+1 overlaps the kernels of step i; every step's copy and its
        # logits read-back are inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        xbuf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
        out_host = torch.empty((batch, NUM_CLASS), dtype=torch.float32).pin_memory()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def e2e_loop(steps):
            with torch.cuda.stream(copy_stream):
                xbuf[0].copy_(x_host, non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(steps):
                cur, nxt = i & 1, (i + 1) & 1
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(consumed[nxt])
                        xbuf[nxt].copy_(x_host, non_blocking=True)
                        ready[nxt].record(copy_stream)
                main.wait_event(ready[cur])
                lg = net(xbuf[cur])
                consumed[cur].record(main)
                out_host.copy_(lg, non_blocking=True)
            main.synchronize()

        note("device-resident timing done")
        e2e_loop(6)                    # each of the two input buffers is seen three times: its forward graph exists before the timed loop
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(args.steps)
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)

        # ---------------- per-layer device times of K1 (roofline of the dominant kernel), separate pass
        k1_ms, rows = 0.0, []
        if rank == 0:
            from fastvideotagging_b200 import ops
            reps = 3

            def conv_flops(L):
                m = L.out_shape[0] * L.out_shape[1] * L.out_shape[2] * L.out_shape[3]
                kk = L.spec.cin * L.spec.kernel[0] * L.spec.kernel[1] * L.spec.kernel[2]
                return m, kk, 2.0 * m * L.spec.cout * kk

            i = 0
            while i < len(plan.layers):                       # one entry per LAUNCH of the inference plan
                L = plan.layers[i]
                B = plan.fused.get(i)
                src = plan._view(L.src)
                if B is not None:                             # K2f: the whole (2+1)D unit in one launch
                    dst = plan._view(B.dst)
                    res = plan._view(B.res) if B.res is not None else None
                    fn = lambda: ops.unit2p1_fwd(L.desc, B.desc, src, L.w_packed, L.scale, L.shift, B.w_packed, B.scale,
                                                 B.shift, res, out=dst)
                    m, kk, fl = conv_flops(L)
                    fl += conv_flops(B)[2]
                    name, ncol, i = L.spec.name + "+" + B.spec.name, L.spec.cout, i + 2
                else:
                    dst = plan._view(L.dst)
                    res = plan._view(L.res) if L.res is not None else None
                    fn = lambda: ops.conv3d_fwd(L.desc, src, L.w_packed, L.scale, L.shift, res, out=dst)
                    m, kk, fl = conv_flops(L)
                    name, ncol, i = L.spec.name, L.spec.cout, i + 1
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                fn()
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                t = a.elapsed_time(b) / reps
                k1_ms += t
                rows.append((name, m, ncol, kk, t, fl / t / 1e9, fl))

    # ---------------- training step (BASELINE configs[2]): fwd + bwd + BCE + NCCL all-reduce + fused SGD
    train = None
    if not args.no_train:
        from fastvideotagging_b200.model import SigmoidBinaryCrossEntropyLoss
        from fastvideotagging_b200.trainer import Trainer
        tb = TRAIN_BATCH_PER_GPU
        net.train()
        trainer = Trainer(net, "sgd", {"learning_rate": 1e-4, "momentum": 0.9, "wd": 1e-4}, kvstore="device")
        if args.input == "u8":
            xt_host = torch.from_numpy(synthetic_clips_u8(tb, seed=7 + rank)).pin_memory()
        else:
            xt_host = torch.from_numpy(synthetic_clips(tb, seed=7 + rank)).pin_memory()
        xt = xt_host.to(dev)
        lab = (torch.rand(tb, NUM_CLASS, device=dev) < 0.03).float()
        lab[:, 0] = 1
        crit = SigmoidBinaryCrossEntropyLoss()

        def train_step(x):
            loss = crit(net(x), lab).mean()
            loss.backward()
            trainer.step(tb * world)
            return loss

        note("inference done, training warm-up")
        for i in range(3):
            train_step(xt)
            torch.cuda.synchronize()
            note("train warm-up step %d done" % i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            last = train_step(xt)
        b.record()
        barrier()
        ms_train = a.elapsed_time(b)
        # end to end: every step copies its clips from pinned host memory (double-buffered on a copy stream, the copy of step
        # i+1 overlaps step i) and reads its loss back to the host
        copy_stream_t = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        tbuf = [torch.empty_like(xt), torch.empty_like(xt)]
        t_ready = [torch.cuda.Event(), torch.cuda.Event()]
        t_used = [torch.cuda.Event(), torch.cuda.Event()]
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def train_e2e_loop(steps):
            with torch.cuda.stream(copy_stream_t):
                tbuf[0].copy_(xt_host, non_blocking=True)
                t_ready[0].record(copy_stream_t)
            for i in range(steps):
                cur, nxt = i & 1, (i + 1) & 1
                if i + 1 < steps:
                    with torch.cuda.stream(copy_stream_t):
                        if i >= 1:
                            copy_stream_t.wait_event(t_used[nxt])
                        tbuf[nxt].copy_(xt_host, non_blocking=True)
                        t_ready[nxt].record(copy_stream_t)
                main.wait_event(t_ready[cur])
                loss = train_step(tbuf[cur])
                t_used[cur].record(main)
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            main.synchronize()

        train_e2e_loop(2)
        barrier()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        train_e2e_loop(args.steps)
        b2.record()
        barrier()
        ms_train_e2e = a2.elapsed_time(b2)
        tplan = list(net._train_plans.values())[0]
        # every rank must hold the same weights after the same number of identical updates: compare an exact (integer) checksum
        csum = net._flat.w.view(torch.int32).to(torch.int64).sum().reshape(1)
        cmin, cmax = csum.clone(), csum.clone()
        if world > 1:
            dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        train = {"ms": ms_train, "ms_e2e": ms_train_e2e, "loss": float(last.item()), "batch": tb,
                 "h2d": xt_host.numel() * xt_host.element_size(), "ranks_identical": bool((cmin == cmax).item())}
        # dominant training kernel: the grouped weight gradient of the 12 conv2_x layers (conv_wgrad_group_kernel, one launch
        # per step — the largest single kernel of the training step), timed alone like the inference kernel above; without
        # grouping (FVT_WGRAD_GROUP=0) the conv2_x 1x3x3 weight gradient (conv_wgrad_slab_kernel, 6 launches/step)
        if rank == 0:
            key = tplan._stage_of[0]
            g = tplan._group_obj.get(key)
            if g is not None:
                fn = g.run
                layers_g = tplan._groups[key]
                wg_flop = sum(2.0 * L.rows * L.cout_real * L.cin_real * L.fwd.kt * L.fwd.kh * L.fwd.kw for L in layers_g)
                wg_label = ("conv_wgrad_group_kernel, weight gradients of the %d stride-1 conv2_x layers (1x3x3 64->144 and 3x1x1 "
                            "144->64) at batch %d in one launch (1 launch/step; the largest kernel of the training step)" % (len(layers_g), tb))
            else:
                L = tplan.layers["comp_0_conv_1_middle"]
                src, dyt = tplan.bufs[L.src], torch.randn(L.out_shape, device=dev).to(torch.bfloat16)
                fn = lambda: tplan._wgrad_now(L, src, dyt)
                wg_flop = 2.0 * L.rows * L.cout_real * L.cin_real * 9
                wg_label = ("conv_wgrad_slab_kernel, conv2_x 1x3x3 64->144 weight gradient at batch %d (6 launches/step; the largest "
                            "kernel share of the training step)" % tb)
            fn()
            torch.cuda.synchronize()
            e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0_.record()
            for _ in range(5):
                fn()
            e1_.record()
            torch.cuda.synchronize()
            wg_ms = e0_.elapsed_time(e1_) / 5
            train["wgrad"] = (wg_ms, wg_flop, wg_label)

    # ---------------- BASELINE configs[3]: 63-tag multi-label heads (LSEP, WARP), Meitu-shape clips 16x112x112,
    # batch 16/GPU, same fwd + bwd + all-reduce + SGD step with the ranking loss kernels in the loop
    c4 = None
    if not args.no_train and not args.no_c4:
        from fastvideotagging_b200.model import LsepLoss, WarpLoss
        from fastvideotagging_b200.trainer import Trainer
        from oracle import r2plus1d as orc
        cb, ct, cc = C4_BATCH_PER_GPU, C4_T, C4_NUM_CLASS
        net4 = R2Plus2D(cc, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=ct // 8).to(dev)
        net4.load_param_dict(orc.randomize_bn(orc.init_params(MODEL_DEPTH, cc, seed=0), seed=1))
        net4.train()
        trainer4 = Trainer(net4, "sgd", {"learning_rate": 1e-4, "momentum": 0.9, "wd": 1e-4}, kvstore="device")
        import numpy as np
        rng = np.random.default_rng(11 + rank)
        x4 = torch.from_numpy(rng.random((cb, 3, ct, HW, HW), dtype=np.float32)).to(dev)
        lab4 = np.zeros((cb, cc), np.float32)
        for r in range(cb):                                   # 1-4 tags per clip, every row keeps negatives
            lab4[r, rng.choice(cc, size=int(rng.integers(1, 5)), replace=False)] = 1
        lab4 = torch.from_numpy(lab4).to(dev)
        c4 = {"batch": cb}
        for tag, crit4 in (("lsep", LsepLoss()), ("warp", WarpLoss(auto_advance=False))):
            state = {"i": 0}

            def c4_step():
                if tag == "warp":                              # global sample index: ranks independent of the GPU count
                    crit4.sample_offset = (state["i"] * world + rank) * cb
                state["i"] += 1
                loss = crit4(net4(x4), lab4).sum()
                loss.backward()
                trainer4.step(cb * world)
                return loss

            for _ in range(3):
                c4_step()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                last4 = c4_step()
            b.record()
            barrier()
            c4[tag] = {"ms": a.elapsed_time(b), "loss": float(last4.item())}
        note("configs[3] done")

    # ---------------- BASELINE configs[4]: ECO-Lite (2D trunk + 3D ResNet head): the 3D head's 3x3x3 residual stages on the
    # stacked 96x16x28x28 trunk features, batch 32, through the same conv kernels.  The reference holds NO code for ECO
    # (model/ECO.py:1-3 is two import lines), so this is throughput + roofline only; parity is unpinned by construction.
    eco = None
    if not args.no_eco:
        from fastvideotagging_b200.model import ECOLite3DHead
        torch.manual_seed(0)
        head = ECOLite3DHead(NUM_CLASS).to(dev).eval()
        eb = 32
        xe = torch.rand(eb, 16, 28, 28, 96, device=dev).to(torch.bfloat16)     # NDHWC bf16, as the 2D trunk would emit
        with torch.no_grad():
            for _ in range(3):
                ye = head(xe)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                ye = head(xe)
            b.record()
            barrier()
        eco = {"ms": a.elapsed_time(b), "batch": eb, "finite": bool(torch.isfinite(ye).all().item()),
               "gflop": ECOLite3DHead.conv_gflop_per_clip(16, 28)}
        if world > 1:
            te = torch.tensor([eco["ms"]], device=dev)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            eco["ms"] = te.item()
        note("configs[4] done")

    # max over ranks
    if world > 1:
        tt = torch.tensor([ms, ms_e2e, train["ms"] if train else 0.0, c4["lsep"]["ms"] if c4 else 0.0,
                           c4["warp"]["ms"] if c4 else 0.0, train["ms_e2e"] if train else 0.0], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt[0].item(), tt[1].item()
        if train:
            train["ms"] = tt[2].item()
            train["ms_e2e"] = tt[5].item()
        if c4:
            c4["lsep"]["ms"], c4["warp"]["ms"] = tt[3].item(), tt[4].item()

    if rank == 0:
        clips = batch * world * args.steps
        value = clips / (ms / 1e3)
        e2e = clips / (ms_e2e / 1e3)
        conv_tflops = batch * GFLOP_PER_CLIP_FWD / k1_ms if k1_ms > 0 else None     # GFLOP / ms = TFLOP/s
        # dominant kernel: unit2p1_fused_kernel — the six conv2_x (2+1)D units (64 -> 144 -> 64), one launch each, the
        # largest single share of the step.  achieved = algorithmic 2*M*N*K of both convolutions of one launch / its mean
        # CUDA-event duration.  (FVT_FUSED_UNIT=0: the unfused conv_slab_fwd_kernel launches of the same layers.)
        dom = [r for r in rows if "+" in r[0] and r[0].startswith("comp_")]
        dom_kernel = "unit2p1_fused_is_kernel, conv2_x unit 1x3x3 64->144 + 3x1x1 144->64"
        # DRAM bytes per launch from `ncu --set full` at batch 48 (profiles/): scaled to this batch
        traffic = FUSED_UNIT_DRAM_BYTES_B48 * batch / 48.0 if FUSED_UNIT_DRAM_BYTES_B48 else None
        if not dom:
            dom = [r for r in rows if r[0].startswith(("comp_0_", "comp_1_", "comp_2_")) and r[0].endswith("_middle")]
            dom_kernel = "conv_slab_fwd_kernel, conv2_x 1x3x3 64->144"
            traffic = 1.952e9 * batch / 48.0          # profiles/r01j_ncu_slab_conv2x_inference.txt
        dom_ms = sum(r[4] for r in dom) / max(len(dom), 1)
        dom_flop = sum(r[6] for r in dom) / max(len(dom), 1)
        dom_tflops = dom_flop / dom_ms / 1e9 if dom else None
        # The kernel is timed ALONE (3 back-to-back launches between two CUDA events), so the denominator is the BURST peak;
        # the sustained figure (what a long step can hold under the power cap) is given beside it.
        roofline = {"bound": "tensor", "achieved": dom_tflops, "peak": peaks["burst"], "unit": "TFLOP/s",
                    "frac": (dom_tflops / peaks["burst"]) if dom_tflops else None,
                    "frac_burst": (dom_tflops / peaks["burst"]) if dom_tflops else None,
                    "frac_sustained": (dom_tflops / peaks["tflops"]) if dom_tflops else None,
                    "peak_sustained": peaks["tflops"], "traffic": traffic,
                    "kernel": "%s (%d launches/step, %.0f%% of the step); "
                              "algorithmic 2MNK = %.1f GFLOP/launch, mean CUDA-event time %.3f ms" % (
                                  dom_kernel, len(dom), 100.0 * dom_ms * len(dom) / (ms / args.steps), dom_flop / 1e9, dom_ms),
                    "peak_source": peaks["source"] + " MEASURED_PEAKS.json: peak = bf16_tflops (burst, best of 10 cuBLAS 8192^3) because the "
                                   "kernel is timed alone; peak_sustained = bf16_tflops_sustained (4 s back to back, power-capped clocks)",
                    "all_conv_kernels": {"launches_per_step": len(rows), "achieved": conv_tflops,
                                         "frac": (conv_tflops / peaks["burst"]) if conv_tflops else None,
                                         "frac_sustained": (conv_tflops / peaks["tflops"]) if conv_tflops else None,
                                         "note": "%d conv launches (unit2p1_fused / conv_igemm_fwd / conv_slab_fwd / "
                                                 "conv_frame_ring / conv_temporal_is), 304.7 GFLOP/clip algorithmic / summed "
                                                 "CUDA-event time" % len(rows)},
                    "step_frac_of_sustained_peak": value * GFLOP_PER_CLIP_FWD / 1e3 / world / peaks["tflops"]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, _, threads = cpu_reference_rate(2, 15, 1)
            cpu = {"value": rate, "unit": "clips/s", "cores": threads, "kind": "port",
                   "sample": "30 clips (15 steps x 2) of the same workload, R34 32x112x112 fp32, oracle torch-CPU restatement "
                             "of model/R2Plus1.py (MXNet itself is not installable), all host threads"}
        line = {
            "metric": "r2plus1d34_32x112_inference_clips_per_s", "value": value, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "R(2+1)D-34 inference, 32x112x112 clips, 101 classes, batch %d/GPU (BASELINE configs[1])" % batch,
                       "global_batch": batch * world, "parallelism": "replicas x%d (batch-sharded, no collective)" % world,
                       "l2": "activations per layer (0.4-1.4 GB) exceed the 126 MB L2; no explicit flush",
                       "gflop_per_clip": GFLOP_PER_CLIP_FWD},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": x_host.numel() * x_host.element_size() * world,
                    "d2h_bytes_per_step": batch * NUM_CLASS * 4 * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": n_launches * args.steps,
            "clocks": clocks,
        }
        if train:
            tv = train["batch"] * world * args.steps / (train["ms"] / 1e3)
            tv_e2e = train["batch"] * world * args.steps / (train["ms_e2e"] / 1e3)
            line["train"] = {"metric": "r2plus1d34_32x112_train_clips_per_s", "value": tv, "unit": "clips/s",
                             "ms_per_step": train["ms"] / args.steps, "final_loss": train["loss"],
                             "config": "BASELINE configs[2]: fwd+bwd, BCE head, batch %d/GPU, SGD-momentum, %s" % (
                                 train["batch"], "NCCL all-reduce bucketed+overlapped" if world > 1 else "single GPU"),
                             "gflop_per_clip": GFLOP_PER_CLIP_TRAIN,
                             "frac_of_tensor_peak": tv * GFLOP_PER_CLIP_TRAIN / 1e3 / world / peaks["tflops"],
                             "e2e": {"value": tv_e2e, "unit": "clips/s", "ms_per_step": train["ms_e2e"] / args.steps,
                                     "h2d_bytes_per_step": train["h2d"] * world, "d2h_bytes_per_step": 4 * world},
                             "ranks_identical": train["ranks_identical"], "deterministic": True}
            if "wgrad" in train:
                wg_ms, wg_flop, wg_label = train["wgrad"]
                wg_tf = wg_flop / wg_ms / 1e9
                line["train"]["roofline"] = {
                    "bound": "tensor", "kernel": wg_label,
                    "achieved": wg_tf, "unit": "TFLOP/s", "peak": peaks["burst"], "frac": wg_tf / peaks["burst"],
                    "frac_burst": wg_tf / peaks["burst"], "frac_sustained": wg_tf / peaks["tflops"],
                    "algorithmic_gflop_per_launch": wg_flop / 1e9, "mean_cuda_event_ms": wg_ms,
                    "note": "timed alone (5 back-to-back launches between two CUDA events): burst peak is the denominator",
                    "step_frac_of_sustained_peak": tv * GFLOP_PER_CLIP_TRAIN / 1e3 / world / peaks["tflops"]}
        if c4:
            line["train_multilabel"] = {
                "config": "BASELINE configs[3]: R(2+1)D-34, 63 tags, 16x112x112 clips, batch %d/GPU, fwd+bwd+SGD" % c4["batch"],
                "gflop_per_clip": GFLOP_PER_CLIP_TRAIN / 2}
            for tag in ("lsep", "warp"):
                v = c4["batch"] * world * args.steps / (c4[tag]["ms"] / 1e3)
                line["train_multilabel"][tag] = {
                    "value": v, "unit": "clips/s", "ms_per_step": c4[tag]["ms"] / args.steps, "final_loss": c4[tag]["loss"],
                    "frac_of_tensor_peak": v * GFLOP_PER_CLIP_TRAIN / 2 / 1e3 / world / peaks["tflops"]}
        if eco:
            ev = eco["batch"] * world * args.steps / (eco["ms"] / 1e3)
            line["eco_lite"] = {
                "config": "BASELINE configs[4]: ECO-Lite 3D head forward (3x3x3 residual stages 96->128->256->512 on stacked "
                          "96x16x28x28 2D-trunk features, i.e. 16x224x224 clips after the 2D trunk), batch %d/GPU, bf16" % eco["batch"],
                "value": ev, "unit": "clips/s", "ms_per_step": eco["ms"] / args.steps, "gflop_per_clip": eco["gflop"],
                "frac_of_tensor_peak": ev * eco["gflop"] / 1e3 / world / peaks["tflops"], "finite": eco["finite"],
                "parity": "unpinned: the reference has no ECO code (model/ECO.py:1-3); tests compare with a torch fp32 restatement of "
                          "the ECO paper's 3D head (tests/test_gpu_heads.py)"}
        emit(line)
        if args.layer_table:
            with open(args.layer_table, "w") as fh:
                fh.write("layer,M,N,K,ms,GFLOP/s\n")
                for r in rows:
                    fh.write("%s,%d,%d,%d,%.4f,%.0f\n" % r[:6])
    if world > 1:
        # Tear down in dependency order: CUDA graphs that captured NCCL kernels first, then the communicator.  A
        # communicator destroyed while such graphs are alive can block forever, so the teardown also has a watchdog.
        sys.stdout.flush()
        net._train_plans.clear()
        net._plans.clear()
        if c4:
            net4._train_plans.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize()
        done = threading.Event()

        def _destroy():
            try:
                dist.destroy_process_group()
            finally:
                done.set()

        threading.Thread(target=_destroy, daemon=True).start()
        if not done.wait(20.0):
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--input", default="u8", choices=["u8", "fp32"],
                    help="clip batches handed to the network: decoded uint8 frames (N,T,H,W,3) or the reference's fp32 (N,3,T,H,W)")
    ap.add_argument("--no-eco", action="store_true", help="skip the configs[4] (ECO-Lite 3D head) measurement")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement")
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[3] (LSEP/WARP, batch 16) training measurement")
    ap.add_argument("--layer-table", default=None, help="write per-layer K1 times (csv)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run as the contract describes
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29541"] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
