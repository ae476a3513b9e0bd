"""Surface-point initialisers (seed points for the region search).

`dichotomy` follows the behaviour of reference backend/main.py:252-326 (the default initialiser and
the one every BASELINE config uses): draw trial points uniformly in a ball, pair points of opposite
sign of f - iso, then bisect every pair until the *mean* |f - iso| over all pairs drops below
`avg_eps` (or `iter_max` bisections).  Differences from the reference, all deliberate:
  * runs on whatever device the model lives on (the reference hard-codes .cuda());
  * takes an explicit `generator` / python `rng` so seeds are reproducible (reference quirk B-11);
  * does not print per-iteration progress unless `verbose`.
"""
import random
import time

import torch


def constraints_filter(x, w_extra_constraints, b_extra_constraints, requires_index=False):
    """Keep points with w_e . x + b_e < 0 for every extra constraint (reference main.py:70-80)."""
    if (w_extra_constraints is not None) and (b_extra_constraints is not None):
        if w_extra_constraints.shape[0] > 0 and x.shape[0] > 0:
            w = w_extra_constraints.to(dtype=x.dtype, device=x.device)
            b = b_extra_constraints.to(dtype=x.dtype, device=x.device)
            index = ((x @ w.t() + b.reshape(1, -1)) < 0).all(dim=1)
            if requires_index:
                return x[index], index
            return x[index]
    if requires_index:
        return x, torch.ones(x.shape[0], dtype=torch.bool, device=x.device)
    return x


def init_within_ball(init_num, init_ball_radius, w_extra_constraints=None, b_extra_constraints=None,
                     generator=None):
    """Rejection-sample `init_num` float32 points inside the ball (reference main.py:83-91)."""
    points = torch.zeros([0, 3], dtype=torch.float32)
    while points.size(0) < init_num:
        p = (torch.rand([init_num - points.size(0), 3], dtype=torch.float32, generator=generator) * 2 - 1) \
            * init_ball_radius
        p = p[(p ** 2).sum(dim=1) < init_ball_radius ** 2]
        p = constraints_filter(p, w_extra_constraints, b_extra_constraints)
        points = torch.cat([points, p], dim=0)
    return points


def _device_of(model):
    for p in model.parameters():
        return p.device
    return torch.device("cpu")


def dichotomy(model, iso, init_num, w_extra_constraints=None, b_extra_constraints=None, try_pts_num=4096,
              init_ball_radius=1.0, iter_max=100, avg_eps=1e-3, time_out=60, provided_surfpts=None,
              provided_surfstd=None, generator=None, rng=None, verbose=False):
    rng = rng if rng is not None else random
    device = _device_of(model)
    t0 = time.time()
    pt_neg = torch.zeros([0, 3], device=device)
    pt_pos = torch.zeros([0, 3], device=device)
    with torch.no_grad():
        while pt_neg.size(0) < init_num:
            if time.time() - t0 > time_out:
                raise Exception(f'Error: dichotomy cannot find solution within {time_out} seconds!')
            if (provided_surfpts is not None) and (provided_surfstd is not None):
                noise = torch.randn(provided_surfpts.shape, generator=generator)
                points = (provided_surfpts + provided_surfstd * noise).to(device)
            else:
                points = init_within_ball(try_pts_num, init_ball_radius, w_extra_constraints,
                                          b_extra_constraints, generator).to(device)
            values = model(points).reshape(-1) - iso
            id_pos = torch.where(values > 0)[0]
            id_neg = torch.where(values < 0)[0]
            n_pos, n_neg = len(id_pos), len(id_neg)
            if n_pos != 0 and n_neg != 0:
                n_range = n_pos * n_neg
                n_sample = min(init_num - pt_neg.size(0), n_range)
                index = torch.tensor(rng.sample(range(n_range), n_sample), device=device)
                pt_pos = torch.cat([pt_pos, points[id_pos[index % n_pos]]], dim=0)
                pt_neg = torch.cat([pt_neg, points[id_neg[torch.div(index, n_pos, rounding_mode='floor')]]], dim=0)
                if verbose:
                    print(f'(cuam) (dichotomy) n_pos = {n_pos} | n_neg = {n_neg}')
        pt_mid = (pt_neg + pt_pos) / 2
        for _ in range(iter_max):
            va_mid = model(pt_mid).reshape(-1) - iso
            avg_err = va_mid.abs().mean()
            if verbose:
                print(f'(cuam) (dichotomy) avg_err = {avg_err}')
            if avg_err < avg_eps:
                break
            neg = va_mid < 0
            pt_neg[neg] = pt_mid[neg]
            pt_pos[~neg] = pt_mid[~neg]
            pt_mid = (pt_neg + pt_pos) / 2
    return pt_mid  # float32, on the model's device


def states_of(model, points):
    """Activation bits (post-ReLU output > 0) of every hidden neuron at `points`, concatenated over
    the hidden layers -> bool (N, L).  Reference backend/main.py:408-411."""
    with torch.no_grad():
        model(points, requires_outputs_list=True)
        states = torch.cat([(s > 0) for s in model.outputs_list], dim=1)
        model.outputs_list = []
    return states
