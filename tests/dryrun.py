"""Dry-run backend for HOST-LOGIC tests on a machine without a GPU (test infrastructure, never on the product path).

The drop-in class has no CPU fallback.  To exercise its host side where the reference lives (this build container: no
GPU) -- dict staging, launch-plan construction, output / loss assembly, checkpoint I/O as driven by the UNMODIFIED
`slam/slam.py` -- the kernel launches are replaced by a stand-in that fills the engine's output buffers with finite
placeholder values.  Nothing numerical is checked through this backend; numerics are the `-m gpu` tests."""
import contextlib

import torch


def _fill(eng):
    """placeholder results in the engine's static buffers: identity poses, positive disparities / depths, finite losses"""
    eng.out_arena.fill_(0.5)
    eng.cam_t_cam.copy_(torch.eye(4).expand_as(eng.cam_t_cam))
    eng.translation.fill_(0.01)
    eng.axis_angle.zero_()
    eng.losses.fill_(0.1)
    eng.losses[19] = 0.0


@contextlib.contextmanager
def dry_run_backend():
    from clslam_b200 import _lib as L
    from clslam_b200 import depth_pose_prediction as dpp
    from clslam_b200 import engine, train_engine
    te = train_engine.TrainEngine
    saved_te = (te.forward, te.forward_backward, te.backward, te.adam)
    saved = (dpp.DepthPosePrediction._require_cuda, engine.Plan.run, engine.AdaptEngine.forward, engine.AdaptEngine.forward_backward,
             engine.AdaptEngine.backward, engine.AdaptEngine.adam, L.check)
    calls = {'forward': 0, 'forward_backward': 0, 'adam': 0, 'plan_runs': 0}

    def count(key, fn=None):
        def inner(self, *a, **k):
            calls[key] += 1
            if fn is not None:
                fn(self)
        return inner

    dpp.DepthPosePrediction._require_cuda = lambda self: L.load()
    engine.Plan.run = count('plan_runs')
    engine.AdaptEngine.forward = count('forward', _fill)
    engine.AdaptEngine.forward_backward = count('forward_backward', _fill)
    engine.AdaptEngine.backward = count('forward_backward')
    engine.AdaptEngine.adam = count('adam')
    te.forward, te.forward_backward, te.backward, te.adam = (count('forward', _fill), count('forward_backward', _fill),
                                                             count('forward_backward'), count('adam'))
    L.check = lambda rc, what='': None          # clslam_set_device fails without a GPU
    dpp.L.check = L.check
    try:
        yield calls
    finally:
        (dpp.DepthPosePrediction._require_cuda, engine.Plan.run, engine.AdaptEngine.forward, engine.AdaptEngine.forward_backward,
         engine.AdaptEngine.backward, engine.AdaptEngine.adam, L.check) = saved
        te.forward, te.forward_backward, te.backward, te.adam = saved_te
        dpp.L.check = L.check
