"""TEST INFRASTRUCTURE (oracle): torch-CPU restatement of the two voxel-grid helpers next to the render path.

    construct_vox_points_closest   models/mvs/mvs_utils.py:536-561 (the `space_min is None` branch run/train_ft.py uses, :141 / :715);
                                   torch_scatter's scatter_mean / scatter_min are restated with index_add_ and a grouped arg-min
    query_vox_grid                 models/neural_points/neural_points.py:814-826, the same torch calls

Only tests/ may import this module.
"""
import torch


def construct_vox_points_closest(xyz_val, vox_res):
    xyz = xyz_val
    xyz_min, xyz_max = torch.min(xyz, dim=-2)[0], torch.max(xyz, dim=-2)[0]
    space_edge = torch.max(xyz_max - xyz_min) * 1.05
    xyz_mid = (xyz_max + xyz_min) / 2
    space_min = xyz_mid - space_edge / 2
    construct_vox_sz = space_edge / vox_res
    xyz_shift = xyz - space_min[None, ...]
    sparse_grid_idx, inv_idx = torch.unique(torch.floor(xyz_shift / construct_vox_sz[None, ...]).to(torch.int32), dim=0, return_inverse=True)
    V = sparse_grid_idx.shape[0]
    cnt = torch.zeros(V).index_add_(0, inv_idx, torch.ones(xyz.shape[0]))
    xyz_centroid = torch.zeros(V, 3).index_add_(0, inv_idx, xyz_val) / cnt[:, None]               # scatter_mean
    xyz_residual = torch.norm(xyz_val - xyz_centroid[inv_idx, :], dim=-1)
    # scatter_min: per voxel the point of smallest residual (the first one on ties)
    order = torch.argsort(xyz_residual, stable=True)
    order = order[torch.argsort(inv_idx[order], stable=True)]
    first = torch.ones(xyz.shape[0], dtype=torch.bool)
    first[1:] = inv_idx[order][1:] != inv_idx[order][:-1]
    min_idx = order[first]
    return xyz_centroid, sparse_grid_idx, min_idx, space_min, construct_vox_sz


def query_vox_grid(sample_loc_w_tensor, full_grid_idx, space_min, grid_vox_sz, grid_res):
    B, R, SR, _ = sample_loc_w_tensor.shape
    vox_ind = torch.floor((sample_loc_w_tensor - space_min[None, None, None, :]) / grid_vox_sz).to(torch.int64)
    shift = torch.as_tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 0], [1, 1, 1]], dtype=torch.int64).reshape(1, 1, 1, 8, 3)
    vox_ind = vox_ind[..., None, :] + shift
    vox_mask = torch.any(torch.logical_or(vox_ind < 0, vox_ind > grid_res).view(B, R, SR, -1), dim=3)
    vox_ind = torch.clamp(vox_ind, min=0, max=grid_res).view(-1, 3)
    inds = full_grid_idx[vox_ind[..., 0], vox_ind[..., 1], vox_ind[..., 2]].view(B, R, SR, 8)
    inds[vox_mask, :] = -1
    inds[torch.any(inds < 0, dim=-1), :] = -1
    return inds.to(torch.int64)
