"""GPU: the `prob == 1` outputs (sgn_probe_outputs) against the torch restatement of
models/neural_points_volumetric_model.py:633-656 (oracle/render_ref.py:probe_outputs).  fp32, |delta| <= 1e-5."""
from types import SimpleNamespace

import pytest
import torch

from oracle import render_ref as rr
from sgnerf_b200 import ops
from tests.test_gpu_aggregate import _random_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,SR,K", [(1, 1, 8), (64, 24, 8), (33, 80, 8), (50, 24, 3)])
def test_probe_outputs_vs_oracle(R, SR, K):
    cfg = rr.agg_config()
    N = 4000
    tables, pidx, loc_w, raydir, campos, rot = _random_case(cfg, N, R, SR, K, seed=31)
    g = torch.Generator().manual_seed(9)
    opacity = torch.rand(1, R, SR, generator=g)
    if R > 4:
        opacity[0, 2] = 0.25                       # a full tie: torch.max returns the first index
        opacity[0, 3, 5:] = opacity[0, 3, 4]       # a partial tie behind a larger value
    weight = torch.rand(1, R, SR, K, generator=g) * (pidx[None] >= 0)
    conf_c = torch.rand(1, R, SR, K, generator=g).clamp(1e-4, 1.0)
    gn = rr.gather_neighbors(tables, pidx[None], rot[None], campos[None])
    ref = rr.probe_outputs(opacity, loc_w[None], weight, conf_c, gn)
    mask = torch.ones(R, dtype=torch.int8)
    if R > 4:
        mask[1] = 0
    out = ops.probe_outputs(opacity[0].cuda(), loc_w.cuda(), pidx.cuda(), weight[0].cuda(), conf_c[0].cuda(), mask.cuda(), tables.xyz.cuda(),
                            tables.embedding.cuda(), tables.color.cuda(), tables.dir.cuda(), tables.conf.cuda())
    hit = mask.bool()
    for k, v in ref.items():
        got = out[k].cpu()
        torch.testing.assert_close(got[hit], v[0][hit].reshape(got[hit].shape), rtol=0, atol=1e-5, msg=k)
        assert float(got[~hit].abs().sum()) == 0.0 if (~hit).any() else True
