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
ENVIRONMENT_ID = -1       # cuam.environment_id() of the environment ENVIRONMENT_STR describes
CONSTRAINTS_JITTER = 1e-8   # reference backend/main.py:381


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def AnalyticMarching(model, save_ply_path='mesh.ply', iso=0.0, scale=1.0, center=[0.0, 0.0, 0.0],
                     w_extra_constraints_=torch.zeros([0, 3]), b_extra_constraints_=torch.zeros([0]),
                     init_configs=DICHOTOMY_DICT, save_polymesh=True, save_float32_verts=True, flip_insideout=False,
                     dtype=torch.float64, voxel_configs=None, seed=None):
    global ENVIRONMENT_STR, ENVIRONMENT_ID
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

    info = model.get_info()
    float_type = 'float64' if dtype == torch.float64 else 'float32'
    nodesnum = model.nodes
    arc_table = info['arc_table'].to(dtype=torch.int32).cpu()
    weights = [w.detach().to(dtype=dtype).contiguous() for w in info['weights']]
    biases = [b.detach().to(dtype=dtype).contiguous() for b in info['biases']]
    arc_tm = [t.detach().to(dtype=dtype, device=device).contiguous() for t in info['arc_tm']]
    w_extra_constraints = w_extra_constraints_.to(dtype=dtype, device=device) + CONSTRAINTS_JITTER
    b_extra_constraints = b_extra_constraints_.to(dtype=dtype, device=device)

    # the environment comes first: the default initialiser runs inside it (on the device)
    environment_str = f"float_type = {float_type}\nnodesnum = {nodesnum}\narc_table = {arc_table}\n" \
                      f"num_extra_constraints = {num_extra_constraints}"
    if cuamlib.environment_id() != ENVIRONMENT_ID:      # somebody called cuam.Init / Destroy directly meanwhile
        ENVIRONMENT_STR = ''
    if environment_str != ENVIRONMENT_STR:
        t_start = time.time()
        if cuamlib.environment_id() != 0:
            cuamlib.Destroy()
        cuamlib.Init(float_type=float_type, nodesnum=nodesnum, arc_table=arc_table,
                     num_extra_constraints=num_extra_constraints)
        _sync()
        return_dict['init_cuda_time'] = time.time() - t_start
        ENVIRONMENT_STR = environment_str
        ENVIRONMENT_ID = cuamlib.environment_id()
    else:
        return_dict['init_cuda_time'] = 0.0

    # ---- surface points + their activation patterns ----
    t_start = time.time()
    native_seeds = (cfg['method'] == 'dichotomy' and args.get('provided_surfpts') is None and voxel_configs is None
                    and os.environ.get('AM_B200_TORCH_INIT', '0') != '1')
    states = points = None
    if native_seeds:
        # csrc/seeds.cuh: sampling, forward passes, pairing, bisection and the packed states all on the device
        rep = cuamlib.seed_dichotomy(weights, biases, arc_tm, w_extra_constraints, b_extra_constraints, iso,
                                     init_num=args['init_num'], try_pts_num=args['try_pts_num'],
                                     init_ball_radius=args['init_ball_radius'], iter_max=args['iter_max'],
                                     avg_eps=args['avg_eps'],
                                     seed=seed if seed is not None else random.getrandbits(63))
        return_dict['init_report'] = rep
    elif cfg['method'] == 'dichotomy':
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
    if points is not None:
        states = states_of(model, points).to(dtype=torch.bool).contiguous()
        points = points.to(dtype=dtype).contiguous()
    _sync()
    return_dict['init_point_time'] = time.time() - t_start
    print(f"(cuam) init_point_time = {return_dict['init_point_time']}")
    print(f"(cuam) init_cuda_time = {return_dict['init_cuda_time']}")

    def check(st):
        # polygons the engine had to drop leave holes and can cut the search off behind them: never silent
        dropped = {k: st[k] for k in ("n_overflow", "n_unbounded", "n_inconsistent") if st[k]}
        if dropped:
            import warnings
            warnings.warn(f"(cuam) {dropped}: some region polygons were dropped (more than 32 vertices / unbounded / "
                          "inconsistent clipping); the mesh may have holes", RuntimeWarning)

    def march(sub_states, sub_points, w_e, b_e, path):
        """one march + stitching; path None: the mesh stays in memory (voxel mode), else it is exported"""
        _sync()
        t0 = time.time()
        cuamlib.AnalyticMarching(weights=weights, biases=biases, states=sub_states, points=sub_points, arc_tm=arc_tm,
                                 w_extra_constraints=w_e.contiguous(), b_extra_constraints=b_e.contiguous(), iso=iso,
                                 flip_insideout=flip_insideout)
        t_am = time.time() - t0   # am_march returns after the device work has completed
        t0 = time.time()
        cuamlib.CombineMesh(scale=scale, center=center)
        mesh = None
        if path is not None:
            cuamlib.ExportMesh(file_path=path, is_polymesh=save_polymesh, is_float32=save_float32_verts)
        else:
            mesh = cuamlib.mesh()
        check(cuamlib.stats())
        return t_am, time.time() - t0, mesh

    if voxel_configs is None:
        return_dict['am_time'], return_dict['export_time'], _ = march(states, points, w_extra_constraints,
                                                                      b_extra_constraints, save_ply_path)
        return_dict['stats'] = cuamlib.stats()
        print(f"(cuam) am_time = {return_dict['am_time']}")
        print(f"(cuam) export_time  = {return_dict['export_time']}")
    else:
        # local-grid mode (reference backend/main.py:475-556): one march per occupied voxel with the voxel's six
        # faces as extra constraints.  The reference writes every voxel to a temporary PLY, parses it back and
        # concatenates python lists (main.py:533-550); here the stitched mesh of a voxel is handed over as flat
        # arrays (am_copy_mesh) and the voxels are concatenated with one index offset each -- no files, no lists.
        from .polymesh import PolyMesh
        voxel_size = voxel_configs['voxel_size']
        index_3d = torch.floor(points / voxel_size).to(dtype=torch.long)
        uniq, inverse = torch.unique(index_3d, dim=0, return_inverse=True)
        am_times, export_times = [], []
        verts, sizes, index, base = [], [], [], 0
        for grid_id in range(uniq.shape[0]):
            sel = torch.where(inverse == grid_id)[0]
            lo = (uniq[grid_id].to(torch.float64) * voxel_size).tolist()
            hi = ((uniq[grid_id].to(torch.float64) + 1) * voxel_size).tolist()
            w_, b_ = get_boundary('cube', min_vert=lo, max_vert=hi)
            w_ = w_.to(dtype=dtype, device=device) + CONSTRAINTS_JITTER
            b_ = b_.to(dtype=dtype, device=device)
            t_am, t_ex, (v, fs, fi) = march(states[sel].contiguous(), points[sel].contiguous(),
                                            torch.cat([w_, w_extra_constraints], dim=0),
                                            torch.cat([b_, b_extra_constraints], dim=0), None)
            am_times.append(t_am)
            export_times.append(t_ex)
            print(f"(cuam) [{grid_id}/{uniq.shape[0]}] am_time = {t_am}")
            verts.append(v)
            sizes.append(fs)
            index.append(fi + base)
            base += len(v)
        t0 = time.time()
        merged = PolyMesh.from_arrays(np.concatenate(verts) if verts else np.zeros((0, 3)),
                                      np.concatenate(sizes) if sizes else np.zeros(0, np.int32),
                                      np.concatenate(index) if index else np.zeros(0, np.int32))
        if not save_polymesh:
            merged.poly2tri()
        merged.save(save_ply_path)
        return_dict['am_time'] = sum(am_times)
        return_dict['export_time'] = sum(export_times) + time.time() - t0
        print(f"(cuam) [total] am_time  = {return_dict['am_time']}")
        print(f"(cuam) [total] export_time  = {return_dict['export_time']}")
    return return_dict


@atexit.register
def when_exit():
    global ENVIRONMENT_STR
    if cuamlib.environment_id() != 0:
        cuamlib.Destroy()
    ENVIRONMENT_STR = ''
