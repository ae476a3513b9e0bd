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


def estimate_am_time(model, reference_engine=False):
    """Estimated am_time (seconds) of `model`: the reference's model  t(l, n) = (a n)^(b l + c) * n^d * l^e * f  for
    l hidden layers of mean width n (reference backend/utils.py:76-90), with the six constants RE-FITTED to this
    engine on one B200 (tools/fit_am_time.py: 21 SAL geometric-init MLPs, depth 2..8, width 64..512, 1024 seeds,
    float64; profiles/r02_fit_am_time.json; largest relative error of the fit 38%).
    reference_engine=True returns the reference's own fit for ITS engine (measured on the cube-clipped 8x512
    workloads: 63-106x slower than this engine, profiles/r02_bench_reference_cuda.json)."""
    if reference_engine:
        a, b, c, d, e, f = 1.94452188, 0.13816182, -0.14536181, 0.59338494, -1.20459825, 1.17841059
    else:
        a, b, c, d, e, f = (0.017110196732240516, 0.1228606661815844, 0.3233928770244429, 1.0322686311872493,
                            1.4752353570366534, 8.491691646153003e-06)
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
