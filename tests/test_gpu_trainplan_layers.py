"""Teacher-forced parity of EVERY distinct (layer shape, dispatched kernel) of the training plan at the BASELINE training
configurations, against the oracle:

    C3 = BASELINE configs[2]: R(2+1)D-34, 4 clips of 32x112x112 per GPU, 101 classes
    C4 = BASELINE configs[3]: R(2+1)D-34, 16 clips of 16x112x112 per GPU, 63 tags

End-to-end gradients of this network at random init are chaotic in bf16 (the oracle's own bf16 and
fp32 runs differ by O(1)), so the north-star's "gradients within rel 1e-2" is checked the way it can be: each layer of
`engine.TrainPlan` is run ALONE, through the very descriptors / packed weights / kernels the plan dispatches, on the
oracle's own tensors for that layer (`oracle.Net.forward(train=True, bf16_storage=True, taps=...)`), and every output
— raw conv output + BatchNorm sums, activated output, BatchNorm backward, weight gradient, data gradient — must match
the oracle's value for the same inputs to 1e-2 of the tensor's max.  Reference: model/R2Plus1.py:19-40, 42-82, 99-114
(forward semantics), MXNet BatchNorm/Convolution backward as restated in oracle/r2plus1d.py.
"""
import numpy as np
import pytest
import torch

from oracle import r2plus1d as orc

pytestmark = pytest.mark.gpu

TOL = 1e-2        # north-star tolerance for the bf16 path, relative to the tensor's max magnitude


def _to_ndhwc(t_ncdhw, c_store, device):
    """oracle tensor (N, C, T, H, W) holding bf16-representable fp32 values -> (N, T, H, W, c_store) bf16 on the GPU."""
    n, c, t, h, w = t_ncdhw.shape
    out = torch.zeros((n, t, h, w, c_store), dtype=torch.bfloat16)
    out[..., :c] = t_ncdhw.permute(0, 2, 3, 4, 1).to(torch.bfloat16)
    return out.to(device)


def _from_ndhwc(t_gpu, c):
    return t_gpu[..., :c].float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def _close(name, what, got, ref, tol=TOL):
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    assert scale > 0, (name, what, "reference is all zero")
    assert err <= tol * scale, "%s %s: max|err| %.4g vs max|ref| %.4g (%.3g > %.3g)" % (name, what, err, scale, err / scale, tol)
    return err / scale


def _bn_backward_ref(raw, dz, gamma, mean, invstd):
    """MXNet BatchNorm backward in training mode (oracle.np_batchnorm_backward restated on torch-CPU fp64 for full-size
    tensors): dx = gamma*inv*(dz - mean(dz) - xhat*mean(dz*xhat)), dgamma = sum(dz*xhat), dbeta = sum(dz)."""
    sh = (1, -1, 1, 1, 1)
    x = raw.double()
    g = dz.double()
    m = float(x.numel() // x.shape[1])
    xhat = (x - mean.double().reshape(sh)) * invstd.double().reshape(sh)
    dbeta = g.sum(dim=(0, 2, 3, 4))
    dgamma = (g * xhat).sum(dim=(0, 2, 3, 4))
    dx = (gamma.double() * invstd.double()).reshape(sh) * (g - dbeta.reshape(sh) / m - xhat * dgamma.reshape(sh) / m)
    return dx.float(), dgamma.float(), dbeta.float()


def _distinct_layers(plan):
    """One representative per distinct (conv geometry, role) of the plan, in forward order, with its block context."""
    seen, out = set(), []
    ctx = {}
    for comp, xin_name, xin_shape, a, b, c, d, sc in plan.blocks:
        for L in (a, b, c, d) + ((sc,) if sc is not None else ()):
            ctx[L.spec.name] = (xin_name, a, b, c, d, sc)
    for name, L in plan.layers.items():
        key = (L.fwd.key(), L.spec.role, L.spec.relu)
        if key in seen:
            continue
        seen.add(key)
        out.append((L, ctx.get(name)))
    return out


def _tap_name_of_buffer(plan, buf_name):
    """Name of the oracle tap that lives in plan.bufs[buf_name] (an activated conv output or a block output)."""
    conv = buf_name[:-4]                                  # "<conv name>:act"
    L = plan.layers[conv]
    if L.spec.role == "temporal_out":                     # block output relu(bn2 + shortcut) is kept in d.act
        return "comp_%d_out" % int(conv.split("_")[1])
    return L.spec.bn


def _oracle_tap_of_buffer(plan, taps, buf_name):
    return taps[_tap_name_of_buffer(plan, buf_name)]


class _Taps(dict):
    """Keeps only the oracle tensors the distinct layers need, stored as bf16 (lossless: with bf16_storage=True every
    tapped tensor the test reads holds bf16-representable values) — the full set would be 25 GB of fp32 at C4."""

    def __init__(self, keep):
        super().__init__()
        self.keep = keep

    def __setitem__(self, k, v):
        if k in self.keep:
            q = v.to(torch.bfloat16)
            assert torch.equal(q.float(), v.float()), k
            super().__setitem__(k, q)

    def __getitem__(self, k):
        return super().__getitem__(k).float()


def _check_config(device, n, t, num_class, seed):
    from fastvideotagging_b200 import ops
    from fastvideotagging_b200.model import R2Plus2D
    depth, hw = 34, 112
    pool = (t // 8, hw // 16, hw // 16)
    params = orc.randomize_bn(orc.init_params(depth, num_class, seed=0), seed=1)
    x = np.random.default_rng(seed).random((n, 3, t, hw, hw), dtype=np.float32)
    # ---- the plan under test (weights packed exactly as a training step packs them)
    net = R2Plus2D(num_class, depth, final_spatial_kernel=pool[1], final_temporal_kernel=pool[0]).to(device)
    net.load_param_dict(params)
    net.train()
    xd = torch.from_numpy(x).to(device)
    plan = net._train_plan(xd)
    plan.refresh_weights(net._weights_signature())
    plan._join_side()
    torch.cuda.synchronize()
    # ---- oracle forward in training mode with bf16 storage emulated; the tensors of the distinct layers are tapped
    keep = set()
    for L, ctx in _distinct_layers(plan):
        keep.add(L.spec.name)
        if L is not plan.stem0:
            keep.add(_tap_name_of_buffer(plan, L.src))
        if L.spec.role == "temporal_out":
            keep.add("comp_%d_out" % int(L.spec.name.split("_")[1]))
            keep.add(_tap_name_of_buffer(plan, ctx[0]))
        elif L.spec.role != "shortcut":
            keep.add(L.spec.bn)
    taps = _Taps(keep)
    ref = orc.Net(params, depth, pool, bf16_storage=True)
    with torch.no_grad():
        ref_logits, _ = ref.forward(x, train=True, taps=taps)
    flat = plan.flat
    gen = torch.Generator().manual_seed(seed + 1)
    report = []
    for L, ctx in _distinct_layers(plan):
        name, spec = L.spec.name, L.spec
        c_out, c_in = L.cout_real, spec.cin
        gname, bname, mname, vname = plan._bn_names(L)
        # ---------------- forward: conv (+ BatchNorm sums) on the oracle's input tensor
        if L is plan.stem0:
            plan.stem.unfold(xd, plan.unfold)
            src = plan.unfold
        else:
            src = _to_ndhwc(_oracle_tap_of_buffer(plan, taps, L.src), L.cin_s, device)
        L.stats.zero_()
        ops.conv3d_fwd(L.fwd, src, L.wp, out=L.raw, stats=L.stats)
        raw_ref = taps[name]
        e_raw = _close(name, "raw conv output", _from_ndhwc(L.raw, c_out), raw_ref)
        assert float(L.raw[..., c_out:].abs().max()) == 0.0 if L.cout_s > c_out else True, (name, "pad channels must stay zero")
        sums = ops.stats_decode(L.stats).cpu()
        rq = _from_ndhwc(L.raw, c_out).double()
        _close(name, "sum x", sums[:c_out].double(), rq.sum(dim=(0, 2, 3, 4)), 1e-3)
        _close(name, "sum x^2", sums[L.cout_s:L.cout_s + c_out].double(), (rq * rq).sum(dim=(0, 2, 3, 4)), 1e-3)
        # ---------------- forward: BatchNorm finalize + apply (+ residual) exactly as the plan's forward does
        run_m = getattr(net, mname).clone()
        run_v = getattr(net, vname).clone()
        args = (L.stats, flat.view(flat.w, gname), flat.view(flat.w, bname), run_m, run_v, L.cout_s, L.rows, plan.eps,
                plan.momentum, L.scale, L.shift, L.mean, L.invstd)
        if spec.role == "shortcut":
            ops.bn_finalize(*args)                          # applied inside the block's last fused pass
            act_ref = None
        elif spec.role == "temporal_out":
            xin_name, a, b, c, d, sc = ctx
            comp = int(name.split("_")[1])
            if sc is not None:                              # shortcut raw + its own BatchNorm constants, teacher-forced
                xin = _to_ndhwc(_oracle_tap_of_buffer(plan, taps, xin_name), sc.cin_s, device)
                sc.stats.zero_()
                ops.conv3d_fwd(sc.fwd, xin, sc.wp, out=sc.raw, stats=sc.stats)
                g2, b2, m2, v2 = plan._bn_names(sc)
                ops.bn_finalize(sc.stats, flat.view(flat.w, g2), flat.view(flat.w, b2), getattr(net, m2).clone(),
                                getattr(net, v2).clone(), sc.cout_s, sc.rows, plan.eps, plan.momentum, sc.scale, sc.shift,
                                sc.mean, sc.invstd)
                ops.bn_finalize_apply(*args, L.raw, L.act, True, res=sc.raw, res_scale=sc.scale, res_shift=sc.shift)
            else:
                xin = _to_ndhwc(_oracle_tap_of_buffer(plan, taps, xin_name), L.cout_s, device)
                ops.bn_finalize_apply(*args, L.raw, L.act, True, res=xin)
            act_ref = taps["comp_%d_out" % comp]
        else:
            ops.bn_finalize_apply(*args, L.raw, L.act, True)
            act_ref = taps[spec.bn]
        mean_ref = raw_ref.double().mean(dim=(0, 2, 3, 4))
        var_ref = raw_ref.double().var(dim=(0, 2, 3, 4), unbiased=False)
        _close(name, "batch mean", L.mean[:c_out].double().cpu(), mean_ref, 2e-3)
        assert torch.allclose(L.invstd[:c_out].double().cpu(), 1.0 / torch.sqrt(var_ref + plan.eps), rtol=5e-3), (name, "inv_std")
        # MXNet running statistics: momentum multiplies the OLD value, biased variance
        assert torch.allclose(run_v.double().cpu(), 0.9 * torch.from_numpy(params[vname]).double() + 0.1 * var_ref, rtol=5e-3, atol=1e-5), (name, "running var")
        e_act = _close(name, "activated output", _from_ndhwc(L.act, c_out), act_ref) if act_ref is not None else 0.0
        # ---------------- backward: BatchNorm backward with the mask form the plan uses for this layer
        dact = torch.zeros(L.out_shape)
        dact[..., :c_out] = torch.randn(L.out_shape[:4] + (c_out,), generator=gen) * 0.05
        dact = dact.to(torch.bfloat16)
        dact_d = dact.to(device)
        raw_q = _from_ndhwc(L.raw, c_out)                    # what the kernels read: OUR bf16 raw output
        g = dact[..., :c_out].float().permute(0, 4, 1, 2, 3)
        sc_t = L.scale[:c_out].cpu().reshape(1, -1, 1, 1, 1)
        sh_t = L.shift[:c_out].cpu().reshape(1, -1, 1, 1, 1)
        draw = torch.empty_like(L.raw)
        off_g = flat.slots[gname][0]
        sums2 = flat.g[off_g:off_g + 2 * L.cout_s]
        if spec.role == "shortcut":
            plan._bn_bwd(L, dact_d, None, draw)
            dz = g
        elif spec.role == "temporal_out":
            dz_out = torch.empty_like(L.raw)
            plan._bn_bwd(L, dact_d, L.act, draw, dz_out=dz_out)
            dz = g * (_from_ndhwc(L.act, c_out) > 0)
            assert torch.equal(_from_ndhwc(dz_out, c_out), dz), (name, "masked block gradient")
        else:
            plan._bn_bwd(L, dact_d, True, draw)             # ReLU mask recomputed from raw*scale + shift
            dz = g * ((raw_q * sc_t + sh_t) > 0)
        gamma = flat.view(flat.w, gname).detach().cpu()
        dx_ref, dg_ref, db_ref = _bn_backward_ref(raw_q, dz, gamma, L.mean[:c_out].cpu(), L.invstd[:c_out].cpu())
        e_bn = _close(name, "BatchNorm backward d(raw)", _from_ndhwc(draw, c_out), dx_ref)
        _close(name, "dgamma", sums2[:c_out].cpu(), dg_ref, 2e-3)
        _close(name, "dbeta", sums2[L.cout_s:L.cout_s + c_out].cpu(), db_ref, 2e-3)
        # ---------------- backward: weight gradient and data gradient of the conv, as the plan launches them
        dy = torch.zeros(L.out_shape)
        dy[..., :c_out] = torch.randn(L.out_shape[:4] + (c_out,), generator=gen) * 0.05
        dy = dy.to(torch.bfloat16)
        dy_d = dy.to(device)
        dy_ref = dy[..., :c_out].float().permute(0, 4, 1, 2, 3).contiguous()
        w_ref = torch.from_numpy(params[L.w_name]).to(torch.bfloat16).float()
        if L is plan.stem0:
            x_ref = torch.from_numpy(x).to(torch.bfloat16).float()
        else:
            x_ref = _oracle_tap_of_buffer(plan, taps, L.src).float()
        with torch.no_grad():
            dw_ref = torch.nn.grad.conv3d_weight(x_ref, w_ref.shape, dy_ref, stride=spec.stride, padding=spec.pad)
        plan._wgrad_now(L, src, dy_d)
        torch.cuda.synchronize()
        e_wg = _close(name, "weight gradient", flat.view(flat.g, L.w_name).detach().cpu(), dw_ref)
        e_dg = 0.0
        if L.need_dgrad:
            with torch.no_grad():
                dx_in_ref = torch.nn.grad.conv3d_input(x_ref.shape, w_ref, dy_ref, stride=spec.stride, padding=spec.pad)
            out = torch.full(L.in_shape, float("nan"), dtype=torch.bfloat16, device=device)
            plan._dgrad(L, dy_d, out)
            torch.cuda.synchronize()
            e_dg = _close(name, "data gradient", _from_ndhwc(out, c_in), dx_in_ref)
            if L.cin_s > c_in:
                assert float(out[..., c_in:].float().abs().max()) == 0.0, (name, "pad channels of the data gradient")
        report.append("%-24s %-13s raw %.1e act %.1e bnbwd %.1e wgrad %.1e dgrad %.1e" % (name, spec.role, e_raw, e_act, e_bn, e_wg, e_dg))
    print("\n".join(report))
    assert len(report) >= 17, "R(2+1)D-34 has 17 distinct conv geometries (incl. the 3 projection shortcuts and the stem)"
    # ---------------- the plan's GROUPED weight-gradient launches (one per residual stage) at these shapes: every layer's
    # result equals the single-layer launch checked against the oracle above, up to fp32 summation order (the group uses
    # fewer pixel splits)
    assert plan._groups, "the plan defers its stride-1 weight gradients to grouped launches"
    for key, layers in plan._groups.items():
        for L in layers:                                  # fresh operands in the plan's own buffers (pad channels are never stored)
            L.draw_own.normal_()
            L.draw_own.mul_(0.05)
            plan.bufs[L.src].normal_()
        plan._run_group(key)
        plan._join_side()
        torch.cuda.synchronize()
        for L in layers:
            got = flat.raw(flat.g, L.w_name).detach().clone()
            single = ops.conv3d_wgrad(L.fwd, plan.bufs[L.src], L.draw_own, torch.empty_like(got), L.cout_real, L.cin_real, ohwi=True)
            torch.cuda.synchronize()
            scale = float(single.abs().max())
            assert scale > 0 and torch.isfinite(got).all(), L.spec.name
            assert float((got - single).abs().max()) <= 1e-4 * scale, (L.spec.name, "grouped vs single weight gradient")     # fp32 order over up to 401408 positions
    return net, plan, params, x, ref_logits


def test_c3_every_trainplan_layer_matches_the_oracle(cuda_device):
    """BASELINE configs[2] shapes: 4 clips of 32x112x112, 101 classes."""
    _check_config(cuda_device, 4, 32, 101, seed=123)


def test_c4_every_trainplan_layer_matches_the_oracle_and_one_full_step(cuda_device):
    """BASELINE configs[3] shapes: 16 clips of 16x112x112, 63 tags — the same per-layer check, then ONE full training
    step with the ranking heads in the loop: the loss values equal the oracle's on the kernel's own logits to 1e-5, and
    the WARP trial counts are bit-exact with the global sample offsets bench.py uses (rank r of world W at step i:
    offset (i*W + r)*batch)."""
    from fastvideotagging_b200.model import LsepLoss, WarpLoss
    from oracle import mlc_loss as oml
    n, t, c = 16, 16, 63
    net, plan, params, x, _ = _check_config(cuda_device, n, t, c, seed=11)
    rng = np.random.default_rng(11)
    lab = np.zeros((n, c), np.float32)
    for r in range(n):                                   # 1-4 tags per clip, every row keeps negatives
        lab[r, rng.choice(c, size=int(rng.integers(1, 5)), replace=False)] = 1
    xd = torch.from_numpy(x).to(cuda_device)
    labd = torch.from_numpy(lab).to(cuda_device)
    logits = net(xd)
    lg = logits.detach().float().cpu().numpy()
    loss = LsepLoss()(logits, labd).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert np.isfinite(net._flat.g.cpu().numpy()).all()
    ref_loss = float(oml.lsep_loss(lg, lab, dtype=np.float64)[0])
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss), (loss.item(), ref_loss)
    # WARP with the offsets of rank 3 of 8 at step 5
    step, world, rank = 5, 8, 3
    crit = WarpLoss(auto_advance=False)
    crit.sample_offset = (step * world + rank) * n
    logits = net(xd)
    lg = logits.detach().float().cpu().numpy()
    wl = crit(logits, labd).sum()
    wl.backward()
    torch.cuda.synchronize()
    L_ref, trials_ref = oml.warp_sample(lg, lab, crit.max_num_trails, oml.rank_weights(crit.label_size), crit.seed, crit.sample_offset)
    ref_wl = float(oml.warp_loss(lg, lab, L_ref, dtype=np.float64)[0])
    assert np.array_equal(crit.last_trials.cpu().numpy(), trials_ref), "sampled WARP trial counts must be bit-exact"
    assert np.array_equal(crit.last_rank.cpu().numpy(), L_ref), "rank weights must be bit-exact"
    assert abs(wl.item() - ref_wl) <= 1e-5 * abs(ref_wl), (wl.item(), ref_wl)
