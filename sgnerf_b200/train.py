"""Training step of the render path without host synchronisation, captured in a CUDA graph.

What it replaces: one iteration of the reference's loop -- NeuralPointsRayMarching.forward on a random-pixel patch, the colour
MSE over the rays that hit the cloud + the zero-one regulariser on conf_coefficient (models/base_rendering_model.py:543-607,
`zero_one_loss_items=conf_coefficient`, weight 1e-4), backward, Adam on the aggregator MLP (lr) and the point tables (plr)
(models/mvs_points_volumetric_model.py:67-109).  The reference compacts the hit rays on the host (`.cpu()` syncs at
query_point_indices_worldcoords.py:834/:946); here rows stay uncompacted and the loss is masked, so nothing in the step depends on
a device value and the whole step -- query, aggregation forward (TF32 tensor-core GEMMs), compositing, loss, backward with the
scatter-add into the point tables, gradient all-reduce, Adam -- is one graph launch.  With world_size > 1 the gradients of all
parameters are summed over ranks in one flat bucket (NCCL) inside the same graph; every rank applies the identical Adam step.
"""
import torch
import torch.distributed as dist

from . import dist as sdist
from . import ops, pipeline


class TrainStep:
    """One training iteration as a replayable unit.  set_inputs() copies a batch into static device buffers, step() runs it.
    With use_graph=True the first step() call runs three ordinary (eager) steps on the current inputs -- they allocate the optimiser
    state and set kernel attributes, and they are real optimiser steps -- then captures the step and replays the capture from then on."""

    def __init__(self, scene, n_rays, near, far, bg_color, lr=5e-4, plr=2e-3, conf_loss_weight=1e-4, precision=ops.PRECISION_TF32,
                 use_graph=True, train_dir=True, group=None, zero_epsilon=1e-3):
        """scene: pipeline.RenderScene (its tensors become the trainable leaves).  n_rays: rays per step on this rank (fixed)."""
        self.scene, self.n_rays, self.near, self.far = scene, int(n_rays), float(near), float(far)
        self.precision, self.conf_w, self.group = precision, float(conf_loss_weight), group
        self.zero_eps = float(zero_epsilon)             # --zero_epsilon of the reference (base_rendering_model.py:119, default 1e-3)
        dev = scene.xyz.device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.net_params = [t.requires_grad_(True) for t in scene.weights + scene.biases]
        self.pt_params = [scene.embedding, scene.color] + ([scene.dirs] if train_dir else []) + ([scene.conf] if scene.conf is not None else [])
        for t in self.pt_params:
            t.requires_grad_(True)
        self.params = self.net_params + self.pt_params
        self.optim = torch.optim.Adam([{"params": self.net_params, "lr": lr}, {"params": self.pt_params, "lr": plr}],
                                      capturable=bool(use_graph), fused=True)
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
        scene.grid()                                   # host-side grid parameters + build, once per cloud version

    # ------------------------------------------------------------------------------------------------
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

    def _body(self):
        q = self.scene.qopt
        out = pipeline.render_rays(self.scene, self.campos, self.camrot, self.raydir, self.near, self.far, self.bg,
                                   precision=self.precision, t=self.t, want_aux=True)
        # the reference's loss (base_rendering_model.py:543-641) on uncompacted rows: colour MSE over the rays that hit the cloud
        # (`ray_masked_coarse_raycolor`, + 1e-6 per colour item) and the zero-one regulariser mean(log(v) + log(1 - v)),
        # v = clamp(conf_coefficient, eps, 1 - eps), over the [R'', SR, K] block of the hit rays.  Sums are normalised by the GLOBAL
        # hit count (all ranks), and gradients are summed over ranks: the multi-GPU step is the single-GPU step on the union of rays.
        hit = (out.ray_mask > 0).float()
        cnt = hit.sum()
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
        self.n_hit.copy_(hit.sum())

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
