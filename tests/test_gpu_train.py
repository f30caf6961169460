"""GPU: the synchronisation-free training step (sgnerf_b200.train.TrainStep) -- loss goes down on a fixed batch, and the CUDA-graph
replay of the step matches the same step launched eagerly (same inputs, same initial state; float atomics in the scatter-add make
the two runs agree to rounding, not bit for bit)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import render_ref as rr
from sgnerf_b200 import ops, pipeline, synth, train
from tests.test_gpu_aggregate import cfg_to_c

pytestmark = pytest.mark.gpu


def _make(precision, use_graph, n_rays=512, n_points=60_000, **kw):
    s = synth.scene_room(n_points, room=(3.0, 3.0, 2.0), width=160, height=120, seed=7)
    tabs = synth.make_point_tables(n_points, 32, 0, seed=0)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=3, bias_scale=0.05)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    scene = pipeline.RenderScene(torch.from_numpy(s.xyz), tabs.embedding.reshape(n_points, -1), tabs.color.reshape(n_points, 3),
                                 tabs.dir.reshape(n_points, 3), tabs.conf.reshape(n_points), [P[n + ".weight"].clone() for n in names],
                                 [P[n + ".bias"].clone() for n in names], cfg_to_c(cfg), pipeline.query_options(), device="cuda")
    ts = train.TrainStep(scene, n_rays, s.near, s.far, torch.ones(3), precision=precision, use_graph=use_graph, **kw)
    g = torch.Generator().manual_seed(1)
    pix = torch.randint(0, s.raydir.shape[0], (n_rays,), generator=g)
    gt = torch.rand(n_rays, 3, generator=g)
    t = pipeline.middle_point_ts(s.near, s.far, 400, "cuda", jitter=0.3, n_rays=n_rays, generator=torch.Generator(device="cuda").manual_seed(2))
    ts.set_inputs(torch.from_numpy(s.campos).cuda(), torch.from_numpy(s.camrotc2w).cuda(), torch.from_numpy(s.raydir)[pix].cuda(), gt.cuda(), t)
    return ts


@pytest.mark.parametrize("precision", [ops.PRECISION_FP32, ops.PRECISION_TF32])
def test_loss_decreases_on_fixed_batch(precision):
    ts = _make(precision, use_graph=False)
    losses = []
    for _ in range(12):
        ts.step()
        losses.append(float(ts.loss))
    assert float(ts.n_hit) > 50
    assert np.isfinite(losses).all() and losses[-1] < 0.8 * losses[0], losses


def test_graph_replay_matches_eager_steps():
    a = _make(ops.PRECISION_TF32, use_graph=False)
    b = _make(ops.PRECISION_TF32, use_graph=True)
    for _ in range(3 + 4):          # the graph object takes three eager warm-up steps before capturing
        a.step()
    for _ in range(4):
        b.step()
    torch.cuda.synchronize()
    assert abs(float(a.loss) - float(b.loss)) < 2e-3 * max(1.0, abs(float(a.loss)))
    # Adam moves a parameter by ~lr per step whatever the size of its gradient, so an element whose gradient is at rounding level
    # may go the other way in the two runs (the scatter-adds are float atomics): compare in the mean, not element by element
    for pa, pb in zip(a.params, b.params):
        assert float((pa.detach() - pb.detach()).abs().mean()) < 2e-4


def test_scene_prune_and_grow_rebuild_the_grid():
    """RenderScene.prune / grow (NeuralPoints.prune / grow_points): the edited cloud renders exactly like a scene built from the same
    tensors from scratch (grid and per-point tables are rebuilt on next use)."""
    n_points = 40_000
    s = synth.scene_room(n_points, room=(3.0, 3.0, 2.0), width=160, height=120, seed=7)
    tabs = synth.make_point_tables(n_points, 32, 0, seed=0, conf_spread=0.3)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=3, bias_scale=0.05)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    mk = lambda xyz, e, c, d, cf: pipeline.RenderScene(xyz, e, c, d, cf, [P[n + ".weight"].clone() for n in names], [P[n + ".bias"].clone() for n in names],
                                                       cfg_to_c(cfg), pipeline.query_options(), device="cuda")
    a = mk(torch.from_numpy(s.xyz), tabs.embedding.reshape(n_points, -1), tabs.color.reshape(n_points, 3), tabs.dir.reshape(n_points, 3),
           tabs.conf.reshape(n_points))
    args = (torch.from_numpy(s.campos).cuda(), torch.from_numpy(s.camrotc2w).cuda(), torch.from_numpy(s.raydir).cuda(), s.near, s.far,
            torch.ones(3, device="cuda"))
    with torch.no_grad():
        before = pipeline.render_rays(a, *args, precision=ops.PRECISION_BF16)
        kept = a.prune(0.9)
        assert 0 < kept < n_points
        g = torch.Generator(device="cuda").manual_seed(1)
        m = 500
        a.grow(a.xyz[:m] + 0.003, torch.rand(m, 32, device="cuda", generator=g) - 0.5, torch.rand(m, 3, device="cuda", generator=g),
               torch.nn.functional.normalize(torch.randn(m, 3, device="cuda", generator=g), dim=-1), torch.ones(m, device="cuda"))
        after = pipeline.render_rays(a, *args, precision=ops.PRECISION_BF16)
        b = mk(a.xyz.clone(), a.embedding.clone(), a.color.clone(), a.dirs.clone(), a.conf.clone())
        fresh = pipeline.render_rays(b, *args, precision=ops.PRECISION_BF16)
    assert torch.equal(after.ray_color, fresh.ray_color) and torch.equal(after.ray_mask, fresh.ray_mask)
    assert not torch.equal(after.ray_color, before.ray_color)


def test_host_frame_renderer_matches_resident_render():
    """pipeline.HostFrameRenderer (pinned host buffers in and out, copies on side streams, double-buffered inputs): three frames with
    three different cameras come back exactly as render_rays renders them from device tensors."""
    n_points = 30_000
    s = synth.scene_room(n_points, room=(3.0, 3.0, 2.0), width=160, height=120, seed=7)
    tabs = synth.make_point_tables(n_points, 32, 0, seed=0)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=3, bias_scale=0.05)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    scene = pipeline.RenderScene(torch.from_numpy(s.xyz), tabs.embedding.reshape(n_points, -1), tabs.color.reshape(n_points, 3),
                                 tabs.dir.reshape(n_points, 3), tabs.conf.reshape(n_points), [P[n + ".weight"] for n in names],
                                 [P[n + ".bias"] for n in names], cfg_to_c(cfg), pipeline.query_options(), device="cuda")
    R = s.raydir.shape[0]
    hfr = pipeline.HostFrameRenderer(scene, R, s.near, s.far, torch.ones(3), precision=ops.PRECISION_BF16)
    cams, outs = [], []
    for k in range(3):
        pos = s.campos + np.array([0.05 * k, -0.03 * k, 0.0], np.float32)
        cams.append(torch.from_numpy(np.concatenate([pos, s.camrotc2w.reshape(-1)]).astype(np.float32)).pin_memory())
        outs.append(torch.empty(R, 3).pin_memory())
    h_ray = torch.from_numpy(s.raydir).pin_memory()
    for k in range(3):
        hfr.render(cams[k], h_ray, outs[k])
    hfr.wait()
    for k in range(3):
        with torch.no_grad():
            ref = pipeline.render_rays(scene, cams[k][:3].cuda(), cams[k][3:].view(3, 3).cuda(), h_ray.cuda(), s.near, s.far,
                                       torch.ones(3, device="cuda"), precision=ops.PRECISION_BF16)
        assert torch.equal(outs[k], ref.ray_color.cpu())
    assert not torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("precision,tol", [(ops.PRECISION_FP32, 1e-3), (ops.PRECISION_BF16, 1e-2)])
def test_full_path_rgb_and_depth_vs_oracle(precision, tol):
    """pipeline.render_rays end to end (query -> aggregation -> fused frame tail) against the oracle's NeuralPointsRayMarching.forward
    restatement: rendered RGB and depth (`coarse_depth`) within 1e-3 with the fp32 kernels (BASELINE.json's bar), within the stated
    1e-2 with the bf16 tensor-core kernels; also the aux path (separate kernels) gives the same depth as the fused tail."""
    from oracle import query_ref as qr
    from tests import util
    s = synth.scene_c0(n_points=20_000, n_rays=400)
    opt = qr.default_opt(SR=24)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=0, bias_scale=0.05)
    tabs = synth.make_point_tables(s.xyz.shape[0], 32, 0, seed=0, conf_spread=0.2)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    scene = pipeline.RenderScene(s.xyz, tabs.embedding, tabs.color, tabs.dir, tabs.conf, [P[n + ".weight"] for n in names],
                                 [P[n + ".bias"] for n in names], ops.agg_cfg(), pipeline.query_options(SR=24), device="cuda")
    t = util.shared_t(s.near, s.far, opt.z_depth_dim)
    args = (torch.from_numpy(s.campos).cuda(), torch.from_numpy(s.camrotc2w).cuda(), torch.from_numpy(s.raydir).cuda(), s.near, s.far,
            torch.ones(3, device="cuda"))
    with torch.no_grad():
        out = pipeline.render_rays(scene, *args, precision=precision, t=t.cuda())
        aux = pipeline.render_rays(scene, *args, precision=precision, t=t.cuda(), want_aux=True)
    o_pidx, o_loc, o_loc_w, o_dirs, o_mask, vsize, _, _ = util.oracle_query(s, opt, t)
    tables = SimpleNamespace(xyz=torch.from_numpy(s.xyz), embedding=tabs.embedding, color=tabs.color, dir=tabs.dir, conf=tabs.conf, label_embedding=None)
    ref = rr.render_from_query(P, cfg, tables, o_pidx, o_loc, o_loc_w, o_dirs, o_mask, torch.from_numpy(s.camrotc2w)[None],
                               torch.from_numpy(s.campos)[None], vsize, torch.ones(3))
    sel = o_mask[0] > 0
    assert int(sel.sum()) > 50
    torch.testing.assert_close(out.ray_color.cpu(), ref.coarse_raycolor[0], rtol=0, atol=tol)
    torch.testing.assert_close(out.depth.cpu()[sel], ref.coarse_depth[0], rtol=0, atol=tol)
    assert float(out.depth.cpu()[~sel].abs().max()) == 0.0 if bool((~sel).any()) else True
    torch.testing.assert_close(aux.depth, out.depth, rtol=0, atol=1e-5)


def test_loss_kernel_matches_the_reference_formula_and_its_autograd():
    """sgn_loss_forward_backward against torch: colour MSE over the hit rays + 1e-6 + w * mean(log v + log(1 - v)), v = clamp(conf, eps, 1 - eps)
    over the [R'', SR, K] block of the hit rays (base_rendering_model.py:543-641), and torch.autograd's gradients of it (zero where the
    clamp is active, zero for the rays that missed)."""
    g = torch.Generator().manual_seed(0)
    R, SR, K = 777, 24, 8
    color = torch.rand(R, 3, generator=g)
    gt = torch.rand(R, 3, generator=g)
    mask = (torch.rand(R, generator=g) < 0.7).to(torch.int8)
    conf = (torch.rand(R, SR, K, generator=g) * 1.2 - 0.1).clamp(1e-4, 1.0)       # values on both sides of [eps, 1 - eps]
    conf[0, 0, :4] = torch.tensor([1e-3, 1.0 - 1e-3, 5e-4, 1.0])
    c_ref, f_ref = color.clone().requires_grad_(True), conf.clone().requires_grad_(True)
    sel = mask > 0
    eps, w = 1e-3, 1e-4
    v = torch.clamp(f_ref[sel], eps, 1 - eps)
    want = torch.nn.functional.mse_loss(c_ref[sel], gt[sel]) + 1e-6 + w * torch.mean(torch.log(v) + torch.log(1 - v))
    want.backward()
    cnt = torch.zeros((), device="cuda")
    ops.loss_hit_count(mask.cuda(), cnt)
    assert float(cnt) == float(sel.sum())
    loss = torch.zeros((), device="cuda")
    d_color, d_conf = ops.loss_forward_backward(color.cuda(), gt.cuda(), mask.cuda(), conf.cuda(), cnt, loss, 1.0, w, eps, 1e-6)
    assert abs(float(loss) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))
    torch.testing.assert_close(d_color.cpu(), c_ref.grad, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(d_conf.cpu(), f_ref.grad, rtol=1e-4, atol=1e-10)
    assert float(d_color.cpu()[~sel].abs().sum()) == 0.0 and float(d_conf.cpu()[~sel].abs().sum()) == 0.0


@pytest.mark.parametrize("C", [32, 3, 1])
def test_adam_rows_equals_dense_torch_adam(C):
    """sgn_adam_rows over several steps with gradients that touch different row subsets (and leave most rows untouched) against
    torch.optim.Adam on the dense table: same parameters (a row that never received a gradient is not visited and does not move; a row
    that did keeps being updated by its decaying moments), gradient rows cleared."""
    g = torch.Generator().manual_seed(1)
    N = 5000
    shape = (N, C) if C > 1 else (N,)
    p0 = torch.randn(*shape, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-3)
    p = p0.clone().cuda()
    grad, m, v = torch.zeros_like(p), torch.zeros_like(p), torch.zeros_like(p)
    active = torch.zeros(N, dtype=torch.uint8, device="cuda")
    step = torch.zeros((), device="cuda")
    for it in range(6):
        rows = torch.randperm(N, generator=g)[:300 + 50 * it]
        gd = torch.zeros(*shape)
        gd[rows] = torch.randn(*((rows.numel(), C) if C > 1 else (rows.numel(),)), generator=g) * (10.0 ** (it - 3))
        ref.grad = gd.clone()
        opt.step()
        grad.copy_(gd.cuda())
        ops.adam_step_count(step)
        ops.adam_rows(p, grad, m, v, active, step, 2e-3)
        assert float(grad.abs().sum()) == 0.0                  # consumed rows are cleared
    torch.testing.assert_close(p.cpu(), ref.detach(), rtol=2e-5, atol=2e-6)
    untouched = active.cpu() == 0
    assert 0.2 < float(untouched.float().mean()) < 0.9 and torch.equal(p.cpu()[untouched], p0[untouched])


def test_direct_step_matches_the_autograd_step():
    """train.TrainStep (library entry points called in order, fused loss kernel, row Adam, flat gradient bucket) against
    train.AutogradTrainStep (torch.autograd over the same kernels, torch loss, dense torch Adam) from the same state on the same batch:
    same hit count, same loss trajectory, same parameters up to the rounding of the float atomics."""
    def build(cls):
        ts = _make(ops.PRECISION_TF32, use_graph=False)
        if cls is train.TrainStep:
            return ts
        sc = ts.scene
        fresh = _make(ops.PRECISION_TF32, use_graph=False)      # same seeds -> same scene / batch; rebuild it as the autograd step
        b = train.AutogradTrainStep(fresh.scene, fresh.n_rays, fresh.near, fresh.far, torch.ones(3), precision=ops.PRECISION_TF32, use_graph=False)
        b.set_inputs(fresh.campos, fresh.camrot, fresh.raydir, fresh.gt, fresh.t)
        return b
    a, b = build(train.TrainStep), build(train.AutogradTrainStep)
    la, lb = [], []
    for _ in range(5):
        a.step(); b.step()
        la.append(float(a.loss)); lb.append(float(b.loss))
    assert float(a.n_hit) == float(b.n_hit) > 50
    assert np.allclose(la, lb, rtol=2e-3, atol=1e-6), (la, lb)
    for pa, pb in zip(a.params, b.params):
        assert float((pa.detach() - pb.detach()).abs().mean()) < 2e-4
    # rows no ray ever touched did not move and were never visited
    emb_active = a.pt_active.bool()
    assert 0 < int(emb_active.sum()) < emb_active.numel()


def test_adam_rows_multi_equals_dense_torch_adam():
    """sgn_adam_rows_multi on the four point tables at once ([N,32], [N,3], [N,3], [N]) against torch.optim.Adam on the dense tables:
    a row is updated when ANY table has (or ever had) a gradient in it; tables without a gradient in that row move by their moments
    only, exactly as dense Adam moves them."""
    g = torch.Generator().manual_seed(2)
    N, Cs = 3001, [32, 3, 3, 1]
    shapes = [(N, c) if c > 1 else (N,) for c in Cs]
    p0 = [torch.randn(*s, generator=g) for s in shapes]
    refs = [p.clone().requires_grad_(True) for p in p0]
    opt = torch.optim.Adam(refs, lr=2e-3)
    ps = [p.clone().cuda() for p in p0]
    grads, ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    active = torch.zeros(N, dtype=torch.uint8, device="cuda")
    step = torch.zeros((), device="cuda")
    for it in range(5):
        for k, (s, c) in enumerate(zip(shapes, Cs)):
            rows = torch.randperm(N, generator=g)[:150 + 40 * it + 10 * k]           # different rows per table
            gd = torch.zeros(*s)
            gd[rows] = torch.randn(*((rows.numel(), c) if c > 1 else (rows.numel(),)), generator=g)
            refs[k].grad = gd.clone()
            grads[k].copy_(gd.cuda())
        opt.step()
        ops.adam_step_count(step)
        ops.adam_rows_multi(ps, grads, ms, vs, active, step, 2e-3)
        assert all(float(x.abs().sum()) == 0.0 for x in grads)
    for p, r in zip(ps, refs):
        torch.testing.assert_close(p.cpu(), r.detach(), rtol=2e-5, atol=2e-6)
    un = active.cpu() == 0
    assert 0.1 < float(un.float().mean()) < 0.9 and all(torch.equal(p.cpu()[un], q[un]) for p, q in zip(ps, p0))


def test_adam_rows_list_is_bit_identical_to_adam_rows_multi():
    """sgn_adam_rows_list (marked rows -> list of active rows -> update of the listed rows only) against sgn_adam_rows_multi (every row's
    gradient read): parameters, both moments, the gradients' clearing and the active flags are equal bit for bit over 6 steps; rows marked
    without a gradient (sample_pidx is a superset of the rows that receive one) stay inactive."""
    g = torch.Generator().manual_seed(5)
    N, Cs = 5003, [32, 3, 3, 1]
    shapes = [(N, c) if c > 1 else (N,) for c in Cs]
    p0 = [torch.randn(*s, generator=g).cuda() for s in shapes]
    A = dict(p=[p.clone() for p in p0], g=[torch.zeros_like(p) for p in p0], m=[torch.zeros_like(p) for p in p0], v=[torch.zeros_like(p) for p in p0],
             active=torch.zeros(N, dtype=torch.uint8, device="cuda"), step=torch.zeros((), device="cuda"))
    B = dict(p=[p.clone() for p in p0], g=[torch.zeros_like(p) for p in p0], m=[torch.zeros_like(p) for p in p0], v=[torch.zeros_like(p) for p in p0],
             active=torch.zeros(N, dtype=torch.uint8, device="cuda"), step=torch.zeros((), device="cuda"),
             lst=torch.zeros(N, dtype=torch.int32, device="cuda"), cnt=torch.zeros(1, dtype=torch.int32, device="cuda"), touched=torch.zeros(N, device="cuda"))
    for it in range(6):
        rows = torch.randperm(N, generator=g)[:300 + 50 * it]
        extra = torch.randperm(N, generator=g)[:200]                                   # marked, but no gradient arrives
        for k, (s, c) in enumerate(zip(shapes, Cs)):
            gd = torch.zeros(*s)
            sub = rows[: rows.numel() - 20 * k]
            gd[sub] = torch.randn(*((sub.numel(), c) if c > 1 else (sub.numel(),)), generator=g)
            A["g"][k].copy_(gd.cuda())
            B["g"][k].copy_(gd.cuda())
        pidx = torch.cat([rows, extra, torch.full((64,), -1, dtype=torch.int64)]).to(torch.int32).cuda().reshape(-1, 2)
        ops.adam_step_count(A["step"])
        ops.adam_rows_multi(A["p"], A["g"], A["m"], A["v"], A["active"], A["step"], 2e-3)
        ops.adam_mark_rows(pidx, B["touched"])
        ops.adam_step_count(B["step"])
        ops.adam_rows_list(B["p"], B["g"], B["m"], B["v"], B["active"], B["lst"], B["cnt"], B["touched"], B["step"], 2e-3)
        for key in ("p", "g", "m", "v"):
            for a, b in zip(A[key], B[key]):
                assert torch.equal(a, b), (it, key)
        assert torch.equal(A["active"], B["active"]) and int(B["cnt"]) == int(B["active"].sum()) and float(B["touched"].abs().sum()) == 0.0
        lst = B["lst"][: int(B["cnt"])].cpu().numpy()
        assert len(set(lst.tolist())) == lst.size and bool(B["active"].cpu().numpy()[lst].all())
    assert 0.1 < float(B["active"].float().mean()) < 0.9


def test_rows_union_pack_unpack_round_trip():
    """sgn_rows_union / sgn_rows_pack (the touched-row gradient exchange): the list is the ascending set of rows with touched != 0, packing
    copies exactly those rows of every table, unpacking puts them back; other rows are never written."""
    g = torch.Generator().manual_seed(9)
    N, Cs = 7001, [32, 3, 3, 1]
    tabs = [torch.randn(N, c, generator=g).cuda() if c > 1 else torch.randn(N, generator=g).cuda() for c in Cs]
    touched = torch.zeros(N, device="cuda")
    rows = torch.randperm(N, generator=g)[:913].cuda()
    touched[rows] = torch.randint(1, 5, (913,), generator=g).float().cuda()           # sums over ranks: any positive value
    lst, cnt = torch.full((N,), -7, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.rows_union(touched, lst, cnt)
    assert int(cnt) == 913 and torch.equal(lst[:913].long(), torch.sort(rows)[0]) and bool((lst[913:] == -7).all())
    stride = 40
    packed = torch.full((N, stride), float("nan"), device="cuda")
    ops.rows_pack(tabs, lst, cnt, packed, stride)
    want = torch.cat([t.reshape(N, -1) for t in tabs], dim=1)[lst[:913].long()]
    assert torch.equal(packed[:913, :39], want) and bool(torch.isnan(packed[913:]).all())
    back = [torch.zeros_like(t) for t in tabs]
    packed[:913, :39] *= 2.0
    ops.rows_pack(back, lst, cnt, packed, stride, unpack=True)
    mask = touched != 0
    for b, t in zip(back, tabs):
        assert torch.equal(b[mask], 2.0 * t[mask]) and float(b[~mask].abs().sum()) == 0.0


@pytest.mark.parametrize("use_graph", [False, True])
def test_sparse_exchange_step_equals_the_dense_step(use_graph):
    """TrainStep with the touched-row exchange forced on a single rank (union -> pack -> [all-reduce] -> unpack around the host read of
    the row count, two CUDA graphs) trains like the plain step: same loss trajectory, same parameters up to the float atomics' rounding."""
    a = _make(ops.PRECISION_TF32, use_graph=use_graph)
    b = _make(ops.PRECISION_TF32, use_graph=use_graph, sparse_exchange="force")
    assert b.sparse and not a.sparse
    la, lb = [], []
    for _ in range(6):
        a.step(); b.step()
        torch.cuda.synchronize()
        la.append(float(a.loss)); lb.append(float(b.loss))
    assert np.allclose(la, lb, rtol=2e-3, atol=1e-6), (la, lb)
    for pa, pb in zip(a.params, b.params):
        assert float((pa.detach() - pb.detach()).abs().mean()) < 2e-4
    assert torch.equal(a.pt_active, b.pt_active)
    n = int(b.x_meta_host[0])
    assert 0 < n < b.scene.xyz.shape[0] // 2 and b.exchange_floats == b.n_net + n * b.x_stride


def test_scene_edit_with_stable_indices_equals_a_fresh_scene():
    """RenderScene.edit (prune -> holes, grow -> fill holes / append, per-point tables updated for the written rows only, grid rebuilt):
    after two rounds of edits the scene renders bit for bit like a scene built from scratch from the same tensors, the rows of the
    surviving points never moved, and the per-point cache equals a freshly built one."""
    n_points = 40_000
    s = synth.scene_room(n_points, room=(3.0, 3.0, 2.0), width=160, height=120, seed=7)
    tabs = synth.make_point_tables(n_points, 32, 0, seed=0, conf_spread=0.3)
    cfg = rr.agg_config()
    P = rr.init_params(cfg, seed=3, bias_scale=0.05)
    names = [n for n, _, _ in rr.layer_shapes(cfg)]
    mk = lambda xyz, e, c, d, cf: pipeline.RenderScene(xyz, e, c, d, cf, [P[n + ".weight"].clone() for n in names], [P[n + ".bias"].clone() for n in names],
                                                       cfg_to_c(cfg), pipeline.query_options(), device="cuda")
    a = mk(torch.from_numpy(s.xyz), tabs.embedding.reshape(n_points, -1), tabs.color.reshape(n_points, 3), tabs.dir.reshape(n_points, 3),
           tabs.conf.reshape(n_points))
    args = (torch.from_numpy(s.campos).cuda(), torch.from_numpy(s.camrotc2w).cuda(), torch.from_numpy(s.raydir).cuda(), s.near, s.far,
            torch.ones(3, device="cuda"))
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        before = pipeline.render_rays(a, *args, precision=ops.PRECISION_BF16)           # builds grid + point cache
        emb0 = a.embedding.clone()
        for rnd, (thr, m) in enumerate(((0.8, 300), (0.85, 15000))):                       # round 2 grows more than it prunes: appends
            new = (a.xyz[a.alive if a.alive is not None else slice(None)][:m] + 0.003, torch.rand(m, 32, device="cuda", generator=g) - 0.5,
                   torch.rand(m, 3, device="cuda", generator=g), torch.nn.functional.normalize(torch.randn(m, 3, device="cuda", generator=g), dim=-1),
                   0.9 + 0.1 * torch.rand(m, device="cuda", generator=g))
            rows = a.edit(prune_thresh=thr, add=new)
            assert rows.numel() == m and torch.equal(a.embedding[rows], new[1])
            after = pipeline.render_rays(a, *args, precision=ops.PRECISION_BF16)
            b = mk(a.xyz.clone(), a.embedding.clone(), a.color.clone(), a.dirs.clone(), a.conf.clone())
            b.alive = a.alive.clone()
            fresh = pipeline.render_rays(b, *args, precision=ops.PRECISION_BF16)
            assert torch.equal(after.ray_color, fresh.ray_color) and torch.equal(after.ray_mask, fresh.ray_mask)
            uncached = pipeline.render_rays(a, *args, precision=ops.PRECISION_BF16, use_point_cache=False)
            assert torch.equal(after.ray_color, uncached.ray_color), "incrementally updated per-point tables differ from tables rebuilt in the call"
        survivors = a.alive[:n_points].clone()
        survivors[rows[rows < n_points]] = False
        survivors &= (a.embedding[:n_points] == emb0).all(-1)
        assert int(survivors.sum()) > 0.5 * n_points                                      # most rows never moved
        assert a.xyz.shape[0] > n_points and not torch.equal(after.ray_color, before.ray_color)
        assert int((~a.alive).sum()) == 0 or float(a.xyz[~a.alive].min()) == pipeline.RenderScene.HOLE


def test_adam_dense_multi_equals_torch_adam():
    """sgn_adam_dense_multi (the MLP's tensors in one launch) against torch.optim.Adam, 5 steps, tensors of 1 .. 70k elements."""
    g = torch.Generator().manual_seed(11)
    shapes = [(256, 284), (256,), (1, 256), (1,), (128, 280), (3, 128), (3,)]
    p0 = [torch.randn(*s, generator=g) for s in shapes]
    refs = [p.clone().requires_grad_(True) for p in p0]
    opt = torch.optim.Adam(refs, lr=5e-4)
    ps = [p.clone().cuda() for p in p0]
    grads, ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    step = torch.zeros((), device="cuda")
    for it in range(5):
        for k, s in enumerate(shapes):
            gd = torch.randn(*s, generator=g) * (0.1 + it)
            refs[k].grad = gd.clone()
            grads[k].copy_(gd.cuda())
        opt.step()
        ops.adam_step_count(step)
        ops.adam_dense_multi(ps, grads, ms, vs, step, 5e-4, zero_grad=True)
        assert all(float(x.abs().sum()) == 0.0 for x in grads)
    for p, r in zip(ps, refs):
        torch.testing.assert_close(p.cpu(), r.detach(), rtol=2e-5, atol=2e-6)
