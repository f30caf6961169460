import numpy as np, torch, sys
sys.path.insert(0, '.')
from oracle import query_pers_ref as qp, query_ref as qr
from sgnerf_b200 import synth
from tests import ref_driver_pers as rd
import tests.test_gpu_pers_query as T
opt = qp.default_opt(max_o=127, P=64)
s = T._scene(400_000, 1024)
hp = qp.get_hyperparameters(opt, s.height, s.width, synth.SCANNET_INTRINSIC, s.near, s.far)
L = rd.lib(8)
import ctypes as C
dev='cuda'
xyz = s.xyz_pers.cuda()[None].contiguous()
dim = [int(v) for v in hp.scaled_vdim]
print("dim", dim, "xyz", xyz.shape, xyz.dtype, xyz[0,:3])
pixel_size, vol = dim[0]*dim[1], dim[0]*dim[1]*dim[2]
f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=dev)
i32 = lambda a: torch.tensor(np.asarray(a, dtype=np.int32), device=dev)
shift, vsize, dim_t = f32(hp.ranges[:3]), f32(hp.scaled_vsize), i32(hp.scaled_vdim)
qs = i32(opt.query_size)
actual = torch.full([1], xyz.shape[1], dtype=torch.int32, device=dev)
coor_occ = torch.zeros([1]+dim, dtype=torch.uint8, device=dev)
counter = torch.zeros([1]+dim, dtype=torch.int8, device=dev)
near_id = torch.full([1,dim[0],dim[1]], dim[2], dtype=torch.int32, device=dev)
far_id = torch.full([1,dim[0],dim[1]], -1, dtype=torch.int32, device=dev)
rc = L.refp_get_occ_vox(rd._p(xyz), rd._p(actual), 1, xyz.shape[1], rd._p(shift), rd._p(vsize), rd._p(dim_t), rd._p(qs), pixel_size, vol, rd._p(coor_occ), rd._p(counter), rd._p(near_id), rd._p(far_id), 0, rd._st())
torch.cuda.synchronize()
print("rc", rc, "occ", int(coor_occ.sum()), "counter<0", int((counter<0).sum()), "far>0", int((far_id>0).sum()))
o = qp.query_uncompacted(opt, hp, s.pixel_idx, s.xyz_pers)
print("oracle mask", int(o[2].sum()), "valid", int((o[0]>=0).sum()))
r = rd.query_grid_point_index(L, s.pixel_idx.cuda()[None], xyz, opt, hp)
print(r[4], r[0].shape)
