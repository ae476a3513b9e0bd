"""``AnalyticMarching(model, ply)`` -- the public entry point.

Same signature, defaults, keyword names, returned timing keys and environment-cache behaviour as
reference backend/main.py:335-559; everything native goes through ``analyticmesh_b200.cuam``
(the C-ABI library).  Differences from the reference, all deliberate:
  * the caller's ``init_configs`` dict is not mutated;
  * ``am_time`` is taken after the march has completed on the device (the reference's timer lacks a
    synchronise, B-12); the dict additionally carries ``stats`` (counters of the engine);
  * seeds can be made reproducible with ``seed=`` (B-11).
"""
import atexit
import copy
import os
import random
import tempfile
import time

import numpy as np
import torch

from . import cuam as cuamlib
from .initializers import dichotomy, states_of
from .utils import get_boundary

SPHERE_TRACING_DICT = {
    'method': 'sphere_tracing',
    'args': {'init_num': 1024, 'step_size_max': 1.0, 'step_size_mul': 0.5, 'step_size_min': 1e-3,
             'init_ball_radius': 1.0, 'avg_eps': 1e-3, 'hist_len': 100, 'corr_min': 0.1, 'time_out': 60},
}
GRADIENT_DESCENT_DICT = {
    'method': 'gradient_descent',
    'args': {'init_num': 1024, 'lr_max': 1e-2, 'lr_mul': 0.3, 'lr_min': 1e-5, 'corr_min': 0.1, 'hist_len': 100,
             'loss_type': 'l1', 'optimizer_type': 'adam', 'accept_bad': True, 'init_ball_radius': 1.0,
             'batch_mul': 2, 'avg_eps': 1e-3, 'time_out': 60},
}
DICHOTOMY_DICT = {
    'method': 'dichotomy',
    'args': {'init_num': 1024, 'try_pts_num': 4096, 'init_ball_radius': 1.0, 'iter_max': 100, 'avg_eps': 1e-3,
             'time_out': 60, 'provided_surfpts': None, 'provided_surfstd': None},
}
VOXEL_DICT = {'voxel_size': 0.1}

ENVIRONMENT_STR = ''
CONSTRAINTS_JITTER = 1e-8   # reference backend/main.py:381


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def AnalyticMarching(model, save_ply_path='mesh.ply', iso=0.0, scale=1.0, center=[0.0, 0.0, 0.0],
                     w_extra_constraints_=torch.zeros([0, 3]), b_extra_constraints_=torch.zeros([0]),
                     init_configs=DICHOTOMY_DICT, save_polymesh=True, save_float32_verts=True, flip_insideout=False,
                     dtype=torch.float64, voxel_configs=None, seed=None):
    global ENVIRONMENT_STR
    assert w_extra_constraints_.shape[0] == b_extra_constraints_.shape[0]
    num_extra_constraints = w_extra_constraints_.shape[0] + (0 if voxel_configs is None else 6)
    device = torch.device('cuda') if torch.cuda.is_available() else None
    if device is None:
        raise RuntimeError("AnalyticMarching needs a CUDA device (there is no CPU fallback)")
    model = model.to(device).eval()
    for p in model.parameters():
        p.requires_grad_(False)

    cfg = copy.deepcopy({'method': init_configs['method'],
                         'args': {k: v for k, v in init_configs['args'].items()
                                  if k not in ('w_extra_constraints', 'b_extra_constraints')}})
    args = cfg['args']
    args['w_extra_constraints'] = w_extra_constraints_
    args['b_extra_constraints'] = b_extra_constraints_
    return_dict = {}

    t_start = time.time()
    if cfg['method'] == 'dichotomy':
        if seed is not None:
            args['generator'] = torch.Generator().manual_seed(seed)
            args['rng'] = random.Random(seed)
        points = dichotomy(model, iso, **args)
    elif cfg['method'] == 'sphere_tracing':
        from .initializers import sphere_tracing
        points = sphere_tracing(model, iso, **args)
    elif cfg['method'] == 'gradient_descent':
        from .initializers import gradient_descent
        points = gradient_descent(model, iso, **args)
    else:
        raise Exception(f"Error: No such {cfg['method']}")
    _sync()
    return_dict['init_point_time'] = time.time() - t_start
    print(f"(cuam) init_point_time = {return_dict['init_point_time']}")

    states = states_of(model, points)
    info = model.get_info()
    float_type = 'float64' if dtype == torch.float64 else 'float32'
    nodesnum = model.nodes
    arc_table = info['arc_table'].to(dtype=torch.int32).cpu()
    weights = [w.detach().to(dtype=dtype).contiguous() for w in info['weights']]
    biases = [b.detach().to(dtype=dtype).contiguous() for b in info['biases']]
    arc_tm = [t.detach().to(dtype=dtype, device=device).contiguous() for t in info['arc_tm']]
    states = states.to(dtype=torch.bool).contiguous()
    points = points.to(dtype=dtype).contiguous()
    w_extra_constraints = w_extra_constraints_.to(dtype=dtype, device=device) + CONSTRAINTS_JITTER
    b_extra_constraints = b_extra_constraints_.to(dtype=dtype, device=device)

    environment_str = f"float_type = {float_type}\nnodesnum = {nodesnum}\narc_table = {arc_table}\n" \
                      f"num_extra_constraints = {num_extra_constraints}"
    if environment_str != ENVIRONMENT_STR:
        t_start = time.time()
        if ENVIRONMENT_STR != '':
            cuamlib.Destroy()
        cuamlib.Init(float_type=float_type, nodesnum=nodesnum, arc_table=arc_table,
                     num_extra_constraints=num_extra_constraints)
        _sync()
        return_dict['init_cuda_time'] = time.time() - t_start
        ENVIRONMENT_STR = environment_str
    else:
        return_dict['init_cuda_time'] = 0.0
    print(f"(cuam) init_cuda_time = {return_dict['init_cuda_time']}")

    def march(sub_states, sub_points, w_e, b_e, path):
        _sync()
        t0 = time.time()
        cuamlib.AnalyticMarching(weights=weights, biases=biases, states=sub_states, points=sub_points, arc_tm=arc_tm,
                                 w_extra_constraints=w_e.contiguous(), b_extra_constraints=b_e.contiguous(), iso=iso,
                                 flip_insideout=flip_insideout)
        t_am = time.time() - t0   # am_march returns after the device work has completed
        t0 = time.time()
        cuamlib.CombineMesh(scale=scale, center=center)
        cuamlib.ExportMesh(file_path=path, is_polymesh=save_polymesh, is_float32=save_float32_verts)
        return t_am, time.time() - t0

    if voxel_configs is None:
        return_dict['am_time'], return_dict['export_time'] = march(states, points, w_extra_constraints,
                                                                   b_extra_constraints, save_ply_path)
        return_dict['stats'] = cuamlib.stats()
        print(f"(cuam) am_time = {return_dict['am_time']}")
        print(f"(cuam) export_time  = {return_dict['export_time']}")
    else:
        # local-grid mode (reference backend/main.py:475-556): one march per occupied voxel with the
        # voxel's six faces as extra constraints, meshes concatenated afterwards.
        from .polymesh import PolyMesh
        voxel_size = voxel_configs['voxel_size']
        index_3d = torch.floor(points / voxel_size).to(dtype=torch.long)
        uniq, inverse = torch.unique(index_3d, dim=0, return_inverse=True)
        am_times, export_times, meshes = [], [], []
        tmp_dir = tempfile.mkdtemp()
        for grid_id in range(uniq.shape[0]):
            sel = torch.where(inverse == grid_id)[0]
            lo = (uniq[grid_id].to(torch.float64) * voxel_size).tolist()
            hi = ((uniq[grid_id].to(torch.float64) + 1) * voxel_size).tolist()
            w_, b_ = get_boundary('cube', min_vert=lo, max_vert=hi)
            w_ = w_.to(dtype=dtype, device=device) + CONSTRAINTS_JITTER
            b_ = b_.to(dtype=dtype, device=device)
            path = os.path.join(tmp_dir, f"{grid_id}.ply")
            t_am, t_ex = march(states[sel].contiguous(), points[sel].contiguous(),
                               torch.cat([w_, w_extra_constraints], dim=0), torch.cat([b_, b_extra_constraints], dim=0),
                               path)
            am_times.append(t_am)
            export_times.append(t_ex)
            print(f"(cuam) [{grid_id}/{uniq.shape[0]}] am_time = {t_am}")
            meshes.append(PolyMesh(path))
            os.remove(path)
        os.rmdir(tmp_dir)
        return_dict['am_time'] = sum(am_times)
        return_dict['export_time'] = sum(export_times)
        verts, faces, base = [], [], 0
        for m in meshes:
            v, f = m.vertices(), m.faces()
            verts.extend(v)
            faces.extend([[i + base for i in face] for face in f])
            base += len(v)
        PolyMesh(vertices=verts, faces=faces, colors=[]).save(save_ply_path)
        print(f"(cuam) [total] am_time  = {return_dict['am_time']}")
        print(f"(cuam) [total] export_time  = {return_dict['export_time']}")
    return return_dict


@atexit.register
def when_exit():
    global ENVIRONMENT_STR
    if ENVIRONMENT_STR != '':
        cuamlib.Destroy()
        ENVIRONMENT_STR = ''
