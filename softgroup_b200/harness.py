"""Harness around the model: host batches in the reference's collate layout, the end-to-end call
(host buffers -> H2D -> GPU point->voxel hashing -> forward -> host results) and the closed-form calibration of
the point-wise heads that gives a synthetic checkpoint a realistic grouping load."""
import numpy as np
import torch

from . import ops

_KEYS = ('coords', 'batch_idxs', 'coords_float', 'feats', 'semantic_labels', 'instance_labels', 'pt_offset_labels')


def to_host_batch(scan, pin=True):
    """numpy scan (softgroup_b200.synth.make_scan) -> dict of CPU tensors, like custom.py:191-256 minus the
    voxel maps (those are produced on the GPU by `run_scan`, or on the CPU by `collate_like_reference`)."""
    b = {}
    for k in _KEYS:
        t = torch.from_numpy(np.ascontiguousarray(scan[k]))
        if k in ('semantic_labels', 'instance_labels'):
            t = t.long()
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        b[k] = t
    b['spatial_shape'] = np.asarray(scan['spatial_shape'])
    b['batch_size'] = int(scan['batch_size'])
    b['scan_ids'] = list(scan['scan_ids'])
    return b


def collate_like_reference(scan):
    """Exactly the reference's dataloader output: voxelization_idx on the CPU (custom.py:239)."""
    b = to_host_batch(scan, pin=False)
    voxel_coords, v2p_map, p2v_map = ops.voxelization_idx(b['coords'], b['batch_size'])
    b.update(voxel_coords=voxel_coords, v2p_map=v2p_map, p2v_map=p2v_map)
    return b


def h2d_bytes(batch):
    return int(sum(v.numel() * v.element_size() for v in batch.values() if isinstance(v, torch.Tensor)))


def pointwise_injection(scan, sigma=0.03, seed=0, logit=8.0):
    """Device-resident synthetic point-wise predictions (see SoftGroup.forward_test `inject_pointwise`)."""
    from . import synth
    scores, off = synth.grouping_inputs(scan, sigma=sigma, seed=seed, logit=logit)
    return torch.from_numpy(scores).cuda(), torch.from_numpy(off).cuda()


def run_scan(model, host_batch, device_only=False, inject_pointwise=None):
    """End-to-end call a user makes: pinned host tensors in, result dict out. Point->voxel hashing runs on the GPU
    (voxelization_idx with CUDA tensors), then SoftGroup.forward_test."""
    dev = {k: (v.cuda(non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    voxel_coords, v2p_map, p2v_map = ops.voxelization_idx(dev['coords'], dev['batch_size'])
    dev.update(voxel_coords=voxel_coords, v2p_map=v2p_map, p2v_map=p2v_map)
    dev.pop('coords')
    return model.forward_test(device_only=device_only, inject_pointwise=inject_pointwise, **dev)


def device_batch(host_batch):
    """Inputs resident in HBM (for the `value` leg of bench.py): everything uploaded and hashed once."""
    dev = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    voxel_coords, v2p_map, p2v_map = ops.voxelization_idx(dev['coords'], dev['batch_size'])
    dev.update(voxel_coords=voxel_coords, v2p_map=v2p_map, p2v_map=p2v_map)
    return dev


@torch.no_grad()
def calibrate_heads(model, host_batch, logit=6.0, ridge=1e-3):
    """Closed-form fit of the LAST Linear of semantic_linear / offset_linear on the (random) backbone features of
    one scan against its synthetic labels (ridge least squares), so that a realistic share of points passes
    score_thr and shifted coordinates collapse towards instance centroids. Returns fit statistics."""
    dev = device_batch(host_batch)
    feats = dev['feats']
    if model.with_coords:
        feats = torch.cat((feats, dev['coords_float']), 1)
    from . import spconv
    vf = ops.voxelization(feats.contiguous(), dev['p2v_map'])
    x = spconv.SparseConvTensor(vf, dev['voxel_coords'].int(), dev['spatial_shape'], dev['batch_size'])
    out = model.output_layer(model.unet(model.input_conv(x))).features
    pf = out[dev['v2p_map'].long()]

    def hidden(mlp):
        h = pf
        for m in list(mlp)[:-1]:
            h = m(h)
        return h

    def fit(mlp, target):
        h = hidden(mlp).double()
        A = torch.cat([h, torch.ones(h.size(0), 1, device=h.device, dtype=h.dtype)], 1)
        G = A.t() @ A + ridge * h.size(0) * torch.eye(A.size(1), device=h.device, dtype=h.dtype)
        sol = torch.linalg.solve(G, A.t() @ target.double())
        mlp[-1].weight.copy_(sol[:-1].t().float())
        mlp[-1].bias.copy_(sol[-1].float())

    sem = dev['semantic_labels']
    C = model.semantic_classes
    onehot = torch.zeros((sem.numel(), C), device=sem.device)
    valid = sem >= 0
    onehot[valid, sem[valid]] = logit
    fit(model.semantic_linear, onehot)
    inst = dev['instance_labels'] >= 0
    off_t = dev['pt_offset_labels'].clone()
    off_t[~inst] = 0
    fit(model.offset_linear, off_t)
    scores = model.semantic_linear(pf)
    offs = model.offset_linear(pf)
    acc = (scores.argmax(1) == sem).float().mean().item()
    res = (offs - off_t)[inst]
    return dict(sem_acc=acc, offset_residual_sigma=res.std().item(), offset_label_sigma=off_t[inst].std().item())
