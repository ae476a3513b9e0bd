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


class _Corr:
    """Correlation of the recent error history with time (reference main.py:94-107): a value near 1
    means the error is still going down steadily; it is used as a divergence / stall detector."""

    def __init__(self, hist_len):
        import numpy as np
        self.np = np
        self.hist_len = hist_len
        self.t = np.arange(hist_len)
        self.hist = []

    def push(self, value):
        if not self.hist:
            self.hist = [value] * self.hist_len
            return 1.0
        self.hist = [value] + self.hist[:-1]
        with self.np.errstate(all="ignore"):
            return float(self.np.corrcoef(self.np.asarray(self.hist), self.t)[0, 1])


def sphere_tracing(model, iso, init_num, step_size_max, step_size_mul, step_size_min, w_extra_constraints,
                   b_extra_constraints, init_ball_radius, avg_eps, hist_len, corr_min, time_out, generator=None,
                   verbose=False):
    """Newton-like marching x <- x - step * f * grad f from random points in the ball, restarting with a
    smaller step when the mean error stops decreasing (behaviour of reference main.py:112-177; suited to
    SDF-like fields).  May return fewer than `init_num` points."""
    import numpy as np
    device = _device_of(model)
    step, t0 = step_size_max, time.time()
    corr = _Corr(hist_len)
    while True:
        pts = init_within_ball(init_num, init_ball_radius, w_extra_constraints, b_extra_constraints, generator).to(device)
        prev_err, done = 1e10, False
        while True:
            pts = pts.detach().requires_grad_(True)
            pred = model(pts) - iso
            err = pred.abs().mean()
            c = corr.push(err.item())
            if verbose:
                print(f'(cuam) (sphere_init) avg_err = {err}, corr = {c}, step_size = {step}')
            if time.time() - t0 > time_out:
                raise Exception(f'Error: sphere_tracing cannot find solution within time={time_out}!')
            if err > prev_err * 2 or not torch.isfinite(err) or err > 100 or c < corr_min or not np.isfinite(c):
                step *= step_size_mul
                if step < step_size_min:
                    raise Exception('Error: sphere_tracing cannot find solution!')
                break
            prev_err = err
            if err < avg_eps:
                done = True
                break
            grad = torch.autograd.grad(pred, pts, grad_outputs=torch.ones_like(pred))[0]
            with torch.no_grad():
                pts = pts - step * pred * grad
        if done:
            break
    pts = pts.detach()
    pts = pts[(pts ** 2).sum(dim=1) < init_ball_radius ** 2]
    return constraints_filter(pts, w_extra_constraints, b_extra_constraints)


def gradient_descent(model, iso, init_num, lr_max, lr_mul, lr_min, corr_min, hist_len, loss_type, optimizer_type,
                     accept_bad, w_extra_constraints, b_extra_constraints, init_ball_radius, batch_mul, avg_eps,
                     time_out, generator=None, verbose=False):
    """Minimise |f - iso| over the point coordinates with Adam/SGD, shrinking the learning rate when the
    error history stalls (behaviour of reference main.py:180-249).  Returns exactly `init_num` points."""
    device = _device_of(model)
    if loss_type == 'smooth_l1':
        scale = avg_eps * 10
        loss_fn = lambda x: (torch.nn.functional.smooth_l1_loss(x / scale, torch.zeros_like(x), reduction='none') * scale).mean()  # noqa: E731
    elif loss_type == 'l1':
        loss_fn = lambda x: x.abs().mean()  # noqa: E731
    elif loss_type == 'l2':
        loss_fn = lambda x: (x ** 2).mean()  # noqa: E731
    else:
        raise Exception(f"Error: No such {loss_type}")
    t0 = time.time()
    found = torch.zeros([0, 3], device=device)
    while found.size(0) < init_num:
        p = init_within_ball(int((init_num - found.size(0)) * batch_mul), init_ball_radius, w_extra_constraints,
                             b_extra_constraints, generator).to(device).requires_grad_()
        if optimizer_type == 'adam':
            opt = torch.optim.Adam([p], lr=lr_max)
        elif optimizer_type == 'sgd':
            opt = torch.optim.SGD([p], lr=lr_max, momentum=0.9)
        else:
            raise Exception(f"Error: No such {optimizer_type}")
        corr = _Corr(hist_len)
        err, pred = 1e9, None
        while err > avg_eps:
            if pred is not None:
                opt.zero_grad()
                loss_fn(pred).backward()
                opt.step()
            pred = model(p) - iso
            err = pred.abs().mean().item()
            c = corr.push(err)
            if verbose:
                print(f'(cuam) (gradient_descent) pred_err_mean = {err:.2e} | corr = {c:.2e}')
            if time.time() - t0 > time_out:
                raise Exception(f'Error: gradient_descent cannot find solution within {time_out} seconds!')
            if c < corr_min:
                corr = _Corr(hist_len)
                for g in opt.param_groups:
                    g['lr'] *= lr_mul
                if opt.param_groups[0]['lr'] < lr_min:
                    break
        if opt.param_groups[0]['lr'] >= lr_min or accept_bad:
            q = constraints_filter(p.detach(), w_extra_constraints, b_extra_constraints)
            found = torch.cat([found, q[:init_num - found.size(0)]], dim=0)
    return found
