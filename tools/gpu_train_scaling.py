"""Diagnostic: training-step time vs per-GPU batch, eager launches vs a captured CUDA graph."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_clips, oracle_params, NUM_CLASS, MODEL_DEPTH, T, HW
from fastvideotagging_b200.model import R2Plus2D, SigmoidBinaryCrossEntropyLoss
from fastvideotagging_b200.trainer import Trainer
dev = torch.device("cuda:0")
params = oracle_params()
for tb in (4, 16):
    net = R2Plus2D(NUM_CLASS, MODEL_DEPTH, final_spatial_kernel=HW // 16, final_temporal_kernel=T // 8).to(dev)
    net.load_param_dict(params); net.train()
    trainer = Trainer(net, "sgd", {"learning_rate": 1e-4, "momentum": 0.9, "wd": 1e-4})
    xt = torch.from_numpy(synthetic_clips(tb, seed=7)).to(dev)
    lab = (torch.rand(tb, NUM_CLASS, device=dev) < 0.03).float(); lab[:, 0] = 1
    crit = SigmoidBinaryCrossEntropyLoss()
    def step():
        loss = crit(net(xt), lab).mean(); loss.backward(); trainer.step(tb); return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): step()
    b.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print("batch %d eager: %.2f ms/step GPU-timeline, CPU issue %.2f ms/step -> %.1f clips/s" % (tb, a.elapsed_time(b) / 10, t_issue * 100, tb * 10 / (a.elapsed_time(b) / 1e3)), flush=True)
    # CUDA graph of the whole step
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2): step()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            static_loss = step()
        torch.cuda.synchronize()
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        a.record()
        for _ in range(10): g.replay()
        b.record(); torch.cuda.synchronize()
        print("batch %d graph: %.2f ms/step -> %.1f clips/s (loss %.4f)" % (tb, a.elapsed_time(b) / 10, tb * 10 / (a.elapsed_time(b) / 1e3), static_loss.item()), flush=True)
    except Exception as e:
        print("graph capture failed:", repr(e)[:300], flush=True)
    del net, trainer
    torch.cuda.empty_cache()
