"""Small helpers of the public API (reference backend/utils.py)."""
import torch


def get_boundary(boundary_type='cube', **kwargs):
    """Plane constraints w.x + b < 0 bounding a primitive; 'cube' needs `min_vert` and `max_vert`.
    Same contract as reference backend/utils.py:5-29."""
    if boundary_type != 'cube':
        raise Exception(f'Error: No such {boundary_type}')
    lo, hi = kwargs['min_vert'], kwargs['max_vert']
    w = torch.cat([torch.eye(3), -torch.eye(3)], dim=0).to(torch.float32)
    b = torch.tensor([-hi[0], -hi[1], -hi[2], lo[0], lo[1], lo[2]], dtype=torch.float32)
    return w, b


def estimate_am_time(model):
    """The reference's fitted wall-time model for ITS engine (reference backend/utils.py:76-90);
    kept for API compatibility -- it does not describe this engine."""
    a, b, c, d, e, f = 1.94452188, 0.13816182, -0.14536181, 0.59338494, -1.20459825, 1.17841059
    nodes = model.nodes
    layers = len(nodes) - 2
    width = sum(nodes[1:-1]) / layers
    return (a * width) ** (b * layers + c) * width ** d * layers ** e * f
