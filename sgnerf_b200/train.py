"""Training step of the render path without host synchronisation, captured in a CUDA graph.

What it replaces: one iteration of the reference's loop -- NeuralPointsRayMarching.forward on a random-pixel patch, the colour
MSE over the rays that hit the cloud + the zero-one regulariser on conf_coefficient (models/base_rendering_model.py:543-641,
`zero_one_loss_items=conf_coefficient`, weight 1e-4, --zero_epsilon 1e-3), backward, Adam on the aggregator MLP (lr) and the point
tables (plr) (models/mvs_points_volumetric_model.py:67-109).  The reference compacts the hit rays on the host (`.cpu()` syncs at
query_point_indices_worldcoords.py:834/:946); here rows stay uncompacted and the loss is masked, so nothing in the step depends on
a device value and the whole step is one graph launch.

TrainStep calls the library's entry points in order -- no autograd graph is built:

    sgn_query -> sgn_agg_forward (save) -> sgn_ray_dist -> sgn_composite_forward -> sgn_loss_hit_count [-> all-reduce of the count]
    -> sgn_loss_forward_backward (loss + d ray_color + d conf_coefficient in one pass) -> sgn_composite_backward -> sgn_agg_backward
    (TF32 tensor-core GEMMs, scatter-add into the point-table gradient accumulators) [-> ONE in-place all-reduce of the flat gradient
    bucket] -> Adam: sgn_adam_dense_multi on the ~30 MLP tensors, sgn_adam_rows_list on the point tables.

All gradients live in one flat fp32 bucket (MLP gradients first, then the point tables') whose slices are the accumulators the backward
adds into, so the multi-GPU exchange is a single NCCL all-reduce of that buffer in place -- no flatten / divide / un-flatten copies:
the loss of every rank is normalised by the GLOBAL hit count, so the sum over ranks IS the gradient of the single-GPU step on the
union of the rays.  sgn_adam_rows updates only the rows that ever received a gradient (identical to dense Adam: the other rows have zero
moments) and clears the gradient rows it consumed, so neither a dense Adam pass over the N-row tables nor a dense memset of their
gradients happens per step.  It is driven by a LIST of the active rows (sgn_adam_rows_list): the rows a step can touch are marked from
sample_pidx, so no kernel reads all N gradient rows to find them.

Several ranks (sparse_exchange, the default): the marks are summed over the ranks right after the query, every rank derives the same
ordered union of touched rows (sgn_rows_union) and packs its gradient rows of that union behind the MLP gradients (sgn_rows_pack);
ONE all-reduce of [MLP gradients | packed rows] replaces the all-reduce of the whole bucket (16 MB instead of 161 MB at C2 on 2 ranks).
The marks are exchanged on a side branch of the step's first CUDA graph while the aggregation forward runs; the size of the union reaches
the host through pinned memory long before the backward ends, so the device never waits for the host (two graphs per step).

AutogradTrainStep is the same iteration written with torch.autograd over the ops' autograd Functions and torch.optim.Adam on every
parameter (the first implementation; kept as the cross-check of TrainStep in tests/test_gpu_train.py).
"""
import torch
import torch.distributed as dist

from . import dist as sdist
from . import ops, pipeline


class _StepBase:
    def __init__(self, scene, n_rays, near, far, bg_color, lr, plr, conf_loss_weight, precision, use_graph, train_dir, group, zero_epsilon,
                 local_only=False):
        self.scene, self.n_rays, self.near, self.far = scene, int(n_rays), float(near), float(far)
        self.precision, self.conf_w, self.group = precision, float(conf_loss_weight), group
        self.zero_eps = float(zero_epsilon)             # --zero_epsilon of the reference (base_rendering_model.py:119, default 1e-3)
        self.lr, self.plr = float(lr), float(plr)
        dev = scene.xyz.device
        self.world = dist.get_world_size(group) if (dist.is_initialized() and not local_only) else 1     # local_only: no exchange (timing aid)
        q = scene.qopt
        # static inputs of the graph: the caller fills them (set_inputs) before every step
        self.raydir = torch.zeros(self.n_rays, 3, device=dev)
        self.gt = torch.zeros(self.n_rays, 3, device=dev)
        self.t = torch.zeros(self.n_rays, q.z_depth_dim, device=dev)
        self.campos = torch.zeros(3, device=dev)
        self.camrot = torch.eye(3, device=dev)
        self.bg = torch.as_tensor(bg_color, dtype=torch.float32, device=dev).reshape(3).clone()
        self.loss = torch.zeros((), device=dev)
        self.n_hit = torch.zeros((), device=dev)
        self.use_graph, self._graph = bool(use_graph), None
        self.train_dir = bool(train_dir)
        scene.grid()                                   # host-side grid parameters + build, once per cloud version

    def set_inputs(self, campos, camrotc2w, raydir, gt, t):
        """Device or host tensors; copied into the graph's static buffers (stream-ordered, no synchronisation)."""
        self.campos.copy_(campos.reshape(3), non_blocking=True)
        self.camrot.copy_(camrotc2w.reshape(3, 3), non_blocking=True)
        self.raydir.copy_(raydir.reshape(-1, 3), non_blocking=True)
        self.gt.copy_(gt.reshape(-1, 3), non_blocking=True)
        self.t.copy_(t, non_blocking=True)

    def jittered_t(self, jitter=0.3, generator=None):
        """Per-ray depth candidates with the reference's training jitter (diff_ray_marching.py:370-386)."""
        return pipeline.middle_point_ts(self.near, self.far, self.scene.qopt.z_depth_dim, self.raydir.device, jitter=jitter,
                                        n_rays=self.n_rays, generator=generator)

    def step(self):
        """One training step on the inputs last given to set_inputs.  Returns nothing; self.loss / self.n_hit are device scalars."""
        self.scene.invalidate_point_cache()      # embeddings and weights are about to change: the bf16 path's per-point tables are stale
        if not self.use_graph:
            self._body()
            return
        if self._graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):                   # eager warm-up: allocates optimizer state, sets kernel attributes
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._body()
        self._graph.replay()


class TrainStep(_StepBase):
    """One training iteration as a replayable unit.  set_inputs() copies a batch into static device buffers, step() runs it.
    With use_graph=True the first step() call runs three ordinary (eager) steps on the current inputs -- they allocate the optimiser
    state and set kernel attributes, and they are real optimiser steps -- then captures the step and replays the capture from then on."""

    def __init__(self, scene, n_rays, near, far, bg_color, lr=5e-4, plr=2e-3, conf_loss_weight=1e-4, precision=ops.PRECISION_TF32,
                 use_graph=True, train_dir=True, group=None, zero_epsilon=1e-3, local_only=False, sparse_exchange=True):
        """scene: pipeline.RenderScene (its tensors are updated in place).  n_rays: rays per step on this rank (fixed).
        sparse_exchange (several ranks): all-reduce the MLP gradients + the point-table gradient rows that some rank touched this step
        instead of the whole bucket; costs one host read of the row count per step (the step becomes two CUDA graphs around it).
        "force" runs the pack / unpack on a single rank too (tests)."""
        super().__init__(scene, n_rays, near, far, bg_color, lr, plr, conf_loss_weight, precision, use_graph, train_dir, group, zero_epsilon,
                         local_only=local_only)
        if precision == ops.PRECISION_BF16:
            raise ValueError("TrainStep runs the fp32 or tf32 path; the bf16 tensor-core path is forward-only")
        dev = scene.xyz.device
        self.net_params = scene.weights + scene.biases
        pts = [("embedding", scene.embedding), ("color", scene.color)] + ([("dirs", scene.dirs)] if train_dir else []) + \
              ([("conf", scene.conf)] if scene.conf is not None else [])
        self.pt_names = [n for n, _ in pts]
        self.pt_params = [t for _, t in pts]
        self.params = self.net_params + self.pt_params
        for t in self.params:
            t.requires_grad_(False)
        # one flat gradient bucket; every parameter's gradient accumulator is a slice of it
        # (every slice starts on a 256-byte boundary: the kernels use 16-byte accesses on the table rows)
        sizes = [p.numel() for p in self.params]
        offs, off = [], 0
        for n in sizes:
            offs.append(off)
            off += (n + 63) // 64 * 64
        n_rows = scene.xyz.shape[0]
        self.flat_grad = torch.zeros(off + n_rows + 64, dtype=torch.float32, device=dev)
        self.grads = [self.flat_grad[o:o + n].view_as(p) for p, n, o in zip(self.params, sizes, offs)]
        # rows this step may touch (1.0 where sample_pidx points): rides at the end of the bucket so the all-reduce unions it over the ranks
        self.pt_touched = self.flat_grad[off:off + n_rows]
        self.sparse = bool(sparse_exchange) and (self.world > 1 or sparse_exchange == "force")
        if self.sparse:
            # exchange buffer [MLP gradients | packed rows]: the MLP accumulators are views of its head, so one all-reduce carries both
            self.x_stride = (sum(p.numel() // n_rows for p in self.pt_params) + 3) // 4 * 4
            self.xbuf = torch.zeros(offs[len(self.net_params)] + n_rows * self.x_stride, dtype=torch.float32, device=dev)
            for i in range(len(self.net_params)):
                self.grads[i] = self.xbuf[offs[i]:offs[i] + sizes[i]].view_as(self.params[i])
            self.x_rows = self.xbuf[offs[len(self.net_params)]:]
            self.x_list = torch.zeros(n_rows, dtype=torch.int32, device=dev)
            self.x_meta = torch.zeros(2, dtype=torch.int32, device=dev)          # [rows in the union, step sequence number]
            self.x_count = self.x_meta[0:1]
            self.x_meta_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._seq = 0
            self._side = torch.cuda.Stream(device=dev)
            self._graph_b = None
        self.exchange_floats = 0 if self.world == 1 else self.flat_grad.numel()       # floats all-reduced by the last step
        self.n_net = offs[len(self.net_params)]           # the MLP gradients occupy flat_grad[:n_net]
        nl = len(scene.weights)
        self.d_w, self.d_b = self.grads[:nl], self.grads[nl:2 * nl]
        self.pt_grads = dict(zip(self.pt_names, self.grads[2 * nl:]))
        # MLP: sgn_adam_dense_multi (all ~30 tensors in one launch; it also clears the accumulators it consumed); point tables:
        # sgn_adam_rows_list with its own moments and active-row list; one step counter for both
        self.net_m = [torch.zeros_like(p) for p in self.net_params]
        self.net_v = [torch.zeros_like(p) for p in self.net_params]
        self.pt_m = [torch.zeros_like(p) for p in self.pt_params]
        self.pt_v = [torch.zeros_like(p) for p in self.pt_params]
        self.pt_active = torch.zeros(scene.xyz.shape[0], dtype=torch.uint8, device=dev)     # rows that ever received a gradient
        self.pt_active_list = torch.zeros(scene.xyz.shape[0], dtype=torch.int32, device=dev)    # ... as a list, in order of first appearance
        self.pt_active_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pt_step = torch.zeros((), dtype=torch.float32, device=dev)
        # number of rays that hit the cloud (the normalisation of both loss terms): with the sparse exchange it sits right behind the
        # touched marks and is summed over the ranks by the same all-reduce
        self._touched_cnt = self.flat_grad[off:off + n_rows + 1]
        self._cnt = self.flat_grad[off + n_rows]

    @torch.no_grad()
    def _body(self):
        self._phase_a()
        if self.sparse:
            self._exchange()
        self._phase_b()

    def step(self):
        if not (self.sparse and self.use_graph):
            return super().step()
        # Sparse exchange: two graphs around the one all-reduce whose size is a host value.  The rows a step can touch are known right
        # after the query: inside the first graph a side branch sums the marks over the ranks, derives the union and copies its size (with
        # a sequence number) to pinned host memory while the main branch runs the aggregation forward.  The host polls that word, then
        # enqueues the right-sized all-reduce and the second graph (unpack + Adam) behind the first: the device never waits for the host.
        self.scene.invalidate_point_cache()
        if self._graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph, self._graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._phase_a()
            with torch.cuda.graph(self._graph_b, pool=self._graph.pool()):
                self._phase_b()
        self._graph.replay()
        self._exchange()
        self._graph_b.replay()

    @torch.no_grad()
    def _exchange(self):
        """Host side of the sparse exchange: wait for this step's row count (a word in pinned memory the device writes early in the step),
        all-reduce that much of the exchange buffer."""
        self._seq += 1
        spins = 0
        while int(self.x_meta_host[1]) != self._seq:
            spins += 1
            if spins > 50_000_000:
                raise RuntimeError("sgnerf_b200.TrainStep: the step's row count never arrived from the device")
        n = int(self.x_meta_host[0])
        self.exchange_floats = self.n_net + n * self.x_stride
        if self.world > 1:
            dist.all_reduce(self.xbuf[:self.exchange_floats], group=self.group)

    @torch.no_grad()
    def _phase_a(self):
        sc, q = self.scene, self.scene.qopt
        grid, hp = sc.grid()
        pidx, loc_w, _, rmask = ops.query(grid, self.campos, self.raydir, self.t, q.SR, q.K, q.kernel_size[0], hp.radius2)
        ops.adam_mark_rows(pidx, self.pt_touched)
        # number of rays that hit the cloud (the normalisation of both loss terms); global after the exchange below
        self._cnt.zero_()
        ops.loss_hit_count(rmask, self._cnt)
        self.n_hit.copy_(self._cnt)
        main = torch.cuda.current_stream()
        if self.sparse:
            # side branch, concurrent with the aggregation forward: marks (+ hit count) summed over the ranks -> union -> its size to the host
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                if self.world > 1:
                    dist.all_reduce(self._touched_cnt, group=self.group)   # 4 B / point: > 0 where some rank has a sample next to the point
                ops.rows_union(self.pt_touched, self.x_list, self.x_count)
                self.x_meta[1:2].add_(1)
                self.x_meta_host.copy_(self.x_meta, non_blocking=True)
        elif self.world > 1:
            dist.all_reduce(self._cnt, group=self.group)
        dec, valid, loc_pers, _, conf, ws, tb = ops.aggregate_train_forward(
            sc.agg_cfg, sc.weights, sc.biases, sc.xyz, sc.embedding, sc.color, sc.dirs, sc.conf, sc.label_emb, pidx, loc_w, self.raydir,
            self.campos, self.camrot, self.precision)
        rd = ops.ray_dist(loc_pers, valid, hp.vsize[2], 1)
        ray_color = ops.composite_forward_raw(dec, rd, valid, self.bg)
        if self.sparse:
            main.wait_stream(self._side)
        d_color, d_conf = ops.loss_forward_backward(ray_color, self.gt, rmask, conf, self._cnt, self.loss, 1.0, self.conf_w, self.zero_eps,
                                                    1e-6 / self.world)
        d_dec = ops.composite_backward_raw(dec, rd, valid, self.bg, d_color)
        # every gradient accumulator is zero here: the Adam kernels clear what they consume
        g = self.pt_grads
        ops.aggregate_train_backward(sc.agg_cfg, sc.weights, sc.biases, tb, pidx, loc_w, self.raydir, self.campos, self.camrot, self.precision,
                                     d_dec, d_conf, self.d_w, self.d_b, g.get("embedding"), g.get("color"), g.get("dirs"), g.get("conf"), ws)
        if self.sparse:
            ops.rows_pack(self.grads[len(self.net_params):], self.x_list, self.x_count, self.x_rows, self.x_stride)
        elif self.world > 1:
            dist.all_reduce(self.flat_grad, group=self.group)          # in place, SUM: losses are normalised by the global hit count

    @torch.no_grad()
    def _phase_b(self):
        if self.sparse:
            ops.rows_pack(self.grads[len(self.net_params):], self.x_list, self.x_count, self.x_rows, self.x_stride, unpack=True)
        ops.adam_step_count(self.pt_step)
        ops.adam_dense_multi(self.net_params, self.grads[:len(self.net_params)], self.net_m, self.net_v, self.pt_step, self.lr, zero_grad=True)
        ops.adam_rows_list(self.pt_params, self.grads[len(self.net_params):], self.pt_m, self.pt_v, self.pt_active, self.pt_active_list,
                           self.pt_active_count, self.pt_touched, self.pt_step, self.plr)


class AutogradTrainStep(_StepBase):
    """The same iteration through torch.autograd (ops._Aggregate / ops._Composite) and torch.optim.Adam on every parameter, dense
    gradients; the multi-GPU exchange flattens the gradients into a bucket and back (dist.allreduce_grads)."""

    def __init__(self, scene, n_rays, near, far, bg_color, lr=5e-4, plr=2e-3, conf_loss_weight=1e-4, precision=ops.PRECISION_TF32,
                 use_graph=True, train_dir=True, group=None, zero_epsilon=1e-3):
        super().__init__(scene, n_rays, near, far, bg_color, lr, plr, conf_loss_weight, precision, use_graph, train_dir, group, zero_epsilon)
        self.net_params = [t.requires_grad_(True) for t in scene.weights + scene.biases]
        self.pt_params = [scene.embedding, scene.color] + ([scene.dirs] if train_dir else []) + ([scene.conf] if scene.conf is not None else [])
        for t in self.pt_params:
            t.requires_grad_(True)
        self.params = self.net_params + self.pt_params
        self.optim = torch.optim.Adam([{"params": self.net_params, "lr": lr}, {"params": self.pt_params, "lr": plr}],
                                      capturable=bool(use_graph), fused=True)

    def _body(self):
        q = self.scene.qopt
        out = pipeline.render_rays(self.scene, self.campos, self.camrot, self.raydir, self.near, self.far, self.bg,
                                   precision=self.precision, t=self.t, want_aux=True)
        # the reference's loss (base_rendering_model.py:543-641) on uncompacted rows, normalised by the GLOBAL hit count
        hit = (out.ray_mask > 0).float()
        cnt = hit.sum()
        self.n_hit.copy_(cnt)
        if self.world > 1:
            dist.all_reduce(cnt, group=self.group)
        cnt = cnt.clamp(min=1.0)
        mse = (((out.ray_color - self.gt) ** 2) * hit[:, None]).sum() / (3.0 * cnt)
        v = out.conf_coef.clamp(self.zero_eps, 1.0 - self.zero_eps)
        zo = ((torch.log(v) + torch.log(1.0 - v)) * hit[:, None, None]).sum() / (cnt * q.SR * q.K)
        loss = mse + 1e-6 / self.world + self.conf_w * zo
        self.optim.zero_grad(set_to_none=True)
        loss.backward()
        if self.world > 1:
            sdist.allreduce_grads(self.params, average=False, group=self.group)
        self.optim.step()
        self.loss.copy_(loss.detach())
