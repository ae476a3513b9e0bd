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


def simplify(ply_path, save_ply_path=None, target_perc=0.01, meshlabserver_path="meshlabserver"):
    """Pass-through to the external `meshlabserver` tool (quadric edge-collapse decimation to `target_perc`
    of the faces, then removal of non-manifold edges and hole closing), like reference backend/utils.py:31-74.
    The tool is not part of this package: a clear error is raised when it is not installed."""
    import os
    import shutil
    import subprocess
    import tempfile
    if shutil.which(meshlabserver_path) is None:
        raise RuntimeError(f"`{meshlabserver_path}` not found: simplify() only drives the external MeshLab server")
    filters = (
        '<!DOCTYPE FilterScript>\n<FilterScript>\n'
        ' <filter name="Simplification: Quadric Edge Collapse Decimation">\n'
        f'  <Param name="TargetPerc" value="{target_perc}" type="RichFloat"/>\n'
        '  <Param name="QualityThr" value="0.5" type="RichFloat"/>\n'
        '  <Param name="PreserveNormal" value="true" type="RichBool"/>\n'
        '  <Param name="OptimalPlacement" value="true" type="RichBool"/>\n'
        '  <Param name="PlanarQuadric" value="true" type="RichBool"/>\n'
        '  <Param name="AutoClean" value="true" type="RichBool"/>\n'
        ' </filter>\n'
        ' <filter name="Select non Manifold Edges "/>\n <filter name="Delete Selected Faces"/>\n'
        ' <filter name="Close Holes">\n  <Param name="MaxHoleSize" value="100" type="RichInt"/>\n </filter>\n'
        '</FilterScript>\n')
    out = ply_path if save_ply_path is None else save_ply_path
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "script.mlx")
        with open(script, "w") as f:
            f.write(filters)
        subprocess.run([meshlabserver_path, "-i", ply_path, "-o", out, "-s", script], check=True)
