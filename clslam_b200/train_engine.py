"""Launch plan of the FULL training step: forward, photometric loss, backward through both
decoders AND both ResNet-18 encoders, with BatchNorm in training mode (batch statistics, running-stat update) or in
evaluation mode with trainable encoder convolutions (``_set_adapt(freeze_encoder=False)``), then Adam over every trainable
tensor.  This is what ``DepthPosePrediction.train()`` (depth_pose_prediction.py:219-289) and ``_set_train()`` (:797-800) need;
the per-frame adaptation step (``AdaptEngine``: frozen encoders, BatchNorm folded into the convolution epilogues, TMA-fed
kernels, three lanes) is a separate, much faster plan.

Design: one lane, fp32 NHWC activations, the generic implicit-GEMM convolution kernels of ``csrc/conv.cu`` (any geometry:
7x7 / stride 2, 3x3 / stride 2, 1x1 / stride 2, reflection-padded 3x3 over an up-sampled + concatenated input) for forward,
input gradients (``clslam_conv_dgrad`` / ``clslam_conv_dgrad_ex``) and weight gradients, BatchNorm / max-pool-backward /
accumulation kernels from ``csrc/train.cu``, and the same photometric / pose-head / loss kernels as the adaptation step.
Offline pre-training is not the latency-critical path: correctness against the reference first, tensor-core tiling later.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib as L
from .engine import Plan, _p, conv_out
from .networks import DEC_CH, ENC_CH, EncoderFlat, FlatParams, ResnetEncoder

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class TrainEngine:
    def __init__(self, *, batch: int, height: int, width: int, batch_global: int, has_sample0: bool,
                 depth_enc: ResnetEncoder, pose_enc: ResnetEncoder, depth_flat: EncoderFlat, pose_flat: EncoderFlat,
                 flat: FlatParams, min_depth: Optional[float], max_depth: Optional[float], disparity_smoothness: float,
                 velocity_loss_scaling: Optional[float], bn_training: bool, device, emit_outputs: bool = True,
                 mask_dynamic: bool = False):
        if height % 32 or width % 32:
            raise ValueError('height and width must be multiples of 32 (ResNet-18 / 4-scale pyramid)')
        self.b, self.h, self.w = batch, height, width
        self.b_global, self.has_sample0 = batch_global, has_sample0
        self.device, self.flat = device, flat
        self.depth_flat, self.pose_flat = depth_flat, pose_flat
        self.bn_training = bn_training
        self.mask_dynamic = mask_dynamic
        if mask_dynamic and (batch_global != batch or batch > 64):
            raise NotImplementedError('mask_dynamic: single process, batch <= 64')
        self.grad_sync_in_plan = False
        self.adam_in_plan = False
        lib = L.load()
        self.lib = lib
        b, h, w = batch, height, width
        plan = Plan(device)
        self.plan = plan
        S = 4
        self.n_scales = S

        # ---- static inputs (the same names as AdaptEngine: the predictor stages into them) -------------------------------
        self.rgb_aug = plan.alloc(3, b, 3, h, w)
        self.target = plan.alloc(b, 3, h, w)
        self.sources = plan.alloc(2, b, 3, h, w)
        self.sources_rgba = plan.alloc(2, b, h, w, 4)
        self.color = [self.target] + [plan.alloc(b, 3, h >> s, w >> s) for s in range(1, S)]
        self.k_mat = plan.alloc(b, 4, 4)
        self.inv_k = plan.alloc(b, 4, 4)
        self.rel_dist = plan.alloc(2, b, dtype=torch.float64, zero=True)
        self.sample_w = plan.alloc(b, zero=True)
        self.sample_w_global = plan.alloc(batch_global, zero=True)
        self.noise_all = plan.alloc(S, b, 2, h, w, zero=True)
        self.noise = [self.noise_all[s] for s in range(S)]
        self.out_views: Dict[tuple, tuple] = {}
        off = 0
        def reserve(key, *shape):
            nonlocal off
            nel = 1
            for d in shape:
                nel *= d
            self.out_views[key] = (off, nel, shape)
            off += (nel + 63) // 64 * 64
        for s in range(S):
            reserve(('disp', s), b, 1, h >> s, w >> s)
        if emit_outputs:
            for s in range(S):
                reserve(('depth', s), b, 1, h, w)
                reserve(('warped', s), 2, b, 3, h, w)
        self.out_arena = plan.alloc(off)

        if mask_dynamic:
            # inputs['mask', 0, s] (:1066, :1081), the weights that turn the reprojection term into the mean over the static pixels,
            # the positions of the b selected gradients per scale, the gradient of the masked smoothness term (csrc/mask.cu)
            self.dyn_mask = [plan.alloc(b, 1, h >> s, w >> s, zero=True) for s in range(S)]
            self.reproj_w = plan.alloc(b, zero=True)
            self.mask_pos = plan.alloc(S, 2, b, dtype=torch.int32, zero=True)
            self.sm_idx = plan.alloc(S, 4 * b, dtype=torch.int32, zero=True)
            self.sm_val = plan.alloc(S, 4 * b, zero=True)
            self.sm_uniform = plan.alloc(S, b, zero=True)
            self._mask_ptrs = (ctypes.c_void_p * 4)(*[_p(t) for t in self.dyn_mask])
            plan.keep.append(self._mask_ptrs)
            plan.simple('mask_prepare', lib.clslam_mask_prepare, ctypes.cast(self._mask_ptrs, ctypes.c_void_p), b, h, w, S, _p(self.reproj_w),
                        _p(self.mask_pos))
        # =================================== forward =====================================================================
        self.identity = plan.alloc(b, 2, h, w)
        plan.simple('pack_rgba', lib.clslam_pack_rgba, _p(self.sources), _p(self.sources_rgba), 2 * b, h, w)
        plan.simple('photo_identity', lib.clslam_photo_identity, _p(self.target), _p(self.sources), _p(self.identity), b, h, w)
        # ---- depth network ------------------------------------------------------------------------------------------------
        self.depth_in = plan.alloc(b, h, w, depth_flat.cin_pad)
        plan.simple('depth.stage', lib.clslam_stage_images, _p(self.rgb_aug[1]), None, _p(self.depth_in), b, h, w, depth_flat.cin_pad)
        self.denc = self._encoder_forward('depth_enc', depth_enc, depth_flat, self.depth_in, b, 1)
        feats = self.denc['feats']
        dd = 'depth_decoder'
        self.y: Dict[tuple, torch.Tensor] = {}
        self.disp: Dict[int, torch.Tensor] = {}
        self.dec_desc: Dict[tuple, tuple] = {}
        x = feats[4]
        for i in range(4, -1, -1):
            hi_, wi_ = x.shape[1], x.shape[2]
            d0 = plan.desc(b, hi_, wi_, x.shape[3], DEC_CH[i], 3, 1, 1, pad_mode=L.PAD_REFLECT, act=L.ACT_ELU)
            y0 = plan.alloc(b, hi_, wi_, DEC_CH[i])
            key0 = (dd, f'upconv_{i}_0.conv.conv')
            plan.conv_fwd(f'upconv_{i}_0', d0, x, None, flat.weight_ptr(key0), None, flat.bias_ptr(key0), None, y0)
            self.y[(i, 0)] = y0; self.dec_desc[(i, 0)] = (d0, x)
            skip = feats[i - 1] if i > 0 else None
            ho_, wo_ = (skip.shape[1], skip.shape[2]) if i > 0 else (hi_ * 2, wi_ * 2)
            c2 = skip.shape[3] if i > 0 else 0
            d1 = plan.desc(b, ho_, wo_, DEC_CH[i] + c2, DEC_CH[i], 3, 1, 1, pad_mode=L.PAD_REFLECT, in_mode=L.IN_UPCAT, c1=DEC_CH[i],
                           h1=hi_, w1=wi_, act=L.ACT_ELU)
            y1 = plan.alloc(b, ho_, wo_, DEC_CH[i])
            key1 = (dd, f'upconv_{i}_1.conv.conv')
            plan.conv_fwd(f'upconv_{i}_1', d1, y0, skip, flat.weight_ptr(key1), None, flat.bias_ptr(key1), None, y1)
            self.y[(i, 1)] = y1; self.dec_desc[(i, 1)] = (d1, y0, skip)
            if i < S:
                keyd = (dd, f'dispconv_{i}.conv')
                dsp = self.out_view(('disp', i))
                plan.simple(f'dispconv_{i}', lib.clslam_dispconv_fwd, _p(y1), flat.weight_ptr(keyd), flat.bias_ptr(keyd), _p(dsp), b, ho_, wo_,
                            DEC_CH[i])
                self.disp[i] = dsp
            x = y1
        # ---- pose network: both passes as one batch of 2b (BatchNorm statistics per pass: groups = 2) ----------------------
        self.pose_in = plan.alloc(2 * b, h, w, pose_flat.cin_pad)
        plan.simple('pose.stage0', lib.clslam_stage_images, _p(self.rgb_aug[0]), _p(self.rgb_aug[1]), _p(self.pose_in[:b]), b, h, w,
                    pose_flat.cin_pad)
        plan.simple('pose.stage1', lib.clslam_stage_images, _p(self.rgb_aug[1]), _p(self.rgb_aug[2]), _p(self.pose_in[b:]), b, h, w,
                    pose_flat.cin_pad)
        self.penc = self._encoder_forward('pose_enc', pose_enc, pose_flat, self.pose_in, 2 * b, 2)
        pf4 = self.penc['feats'][4]
        hp, wp = pf4.shape[1], pf4.shape[2]
        pd = 'pose_decoder'
        self.pose_act, self.pose_desc = {}, {}
        x = pf4
        for name, co, ci, k, act in (('squeeze', 256, 512, 1, L.ACT_RELU), ('pose_0', 256, 256, 3, L.ACT_RELU),
                                     ('pose_1', 256, 256, 3, L.ACT_RELU), ('pose_2', 12, 256, 1, L.ACT_NONE)):
            d = plan.desc(2 * b, hp, wp, ci, co, k, 1, (k - 1) // 2, act=act)
            yv = plan.alloc(2 * b, hp, wp, co)
            plan.conv_fwd(name, d, x, None, flat.weight_ptr((pd, name)), None, flat.bias_ptr((pd, name)), None, yv)
            self.pose_act[name] = yv; self.pose_desc[name] = (d, x)
            x = yv
        self.axis_angle = plan.alloc(2, b, 3)
        self.translation = plan.alloc(2, b, 3)
        self.cam_t_cam = plan.alloc(2, b, 4, 4)
        self.proj = plan.alloc(2, b, 12)
        self.invert_mask = 0b01
        plan.simple('pose_head_fwd', lib.clslam_pose_head_fwd, _p(self.pose_act['pose_2']), _p(self.k_mat), b, hp, wp, self.invert_mask,
                    _p(self.axis_angle), _p(self.translation), _p(self.cam_t_cam), _p(self.proj))
        # ---- photometric chain + losses (same kernels as the adaptation step) -------------------------------------------
        self.n_tiles = lib.clslam_photo_tiles(h, w)
        self.depth = [self.out_view(('depth', s)) for s in range(S)] if emit_outputs else [None] * S
        self.warped = [self.out_view(('warped', s)) for s in range(S)] if emit_outputs else [None] * S
        self.argmin = [plan.alloc(b, h, w, dtype=torch.uint8) for _ in range(S)]
        self.loss_partials = plan.alloc(S, b, self.n_tiles)
        self.ddisp_full = [plan.alloc(b, h, w) for _ in range(S)]
        self.dproj_partials = plan.alloc(S * self.n_tiles, 2, b, 12)
        self.losses = plan.alloc(L.N_LOSS_SLOTS, zero=True)
        self.smooth_grad = plan.alloc(S, 2 * (batch_global + 1) + 1, zero=True)
        if min_depth is None and max_depth is None:
            mode = L.DEPTH_INV
        elif max_depth is None:
            mode = L.DEPTH_MIN_ONLY
        elif min_depth is None:
            raise ValueError('min_depth is None')
        else:
            mode = L.DEPTH_MIN_MAX
        pa = L.PhotoArgs()
        pa.b, pa.h, pa.w, pa.n_scales = b, h, w, S
        pa.depth_mode, pa.min_depth, pa.max_depth = mode, float(min_depth or 0.0), float(max_depth or 0.0)
        pa.target, pa.sources, pa.inv_k, pa.proj = _p(self.target), _p(self.sources), _p(self.inv_k), _p(self.proj)
        pa.sources_rgba = _p(self.sources_rgba)
        pa.identity, pa.sample_w = _p(self.identity), _p(self.sample_w)
        for s in range(S):
            pa.disp[s] = _p(self.disp[s]); pa.noise[s] = _p(self.noise[s])
            pa.depth[s] = _p(self.depth[s]); pa.warped[s] = _p(self.warped[s])
            pa.argmin[s] = _p(self.argmin[s]); pa.ddisp_full[s] = _p(self.ddisp_full[s])
        pa.loss_partials, pa.dproj_partials = _p(self.loss_partials), _p(self.dproj_partials)
        if mask_dynamic:
            pa.dyn_mask, pa.reproj_w = _p(self.dyn_mask[0]), _p(self.reproj_w)
        self.photo_args = pa
        plan.keep.append(pa)
        self.vel_scale = float(velocity_loss_scaling) if (velocity_loss_scaling is not None and velocity_loss_scaling > 0) else 0.0
        la = L.LossArgs()
        la.b, la.b_global, la.h, la.w, la.n_scales, la.n_tiles = b, batch_global, h, w, S, self.n_tiles
        la.has_sample0 = 1 if (has_sample0 and not mask_dynamic) else 0          # masked: the smoothness term comes from csrc/mask.cu
        if mask_dynamic:
            la.reproj_w = _p(self.reproj_w)
        la.disparity_smoothness, la.velocity_scale = float(disparity_smoothness), self.vel_scale
        la.loss_partials, la.sample_w, la.sample_w_global = _p(self.loss_partials), _p(self.sample_w), _p(self.sample_w_global)
        for s in range(S):
            la.disp[s] = _p(self.disp[s]); la.color[s] = _p(self.color[s])
        la.translation, la.rel_dist = _p(self.translation), _p(self.rel_dist)
        la.losses, la.smooth_grad = _p(self.losses), _p(self.smooth_grad)
        self.loss_args = la
        plan.keep.append(la)
        plan.add('photo_fwd', lambda stream, ws: lib.clslam_photo_fwd(ctypes.byref(pa), stream))
        plan.add('loss_finalize', lambda stream, ws: lib.clslam_loss_finalize(ctypes.byref(la), stream))
        if mask_dynamic:
            self._disp_ptrs = (ctypes.c_void_p * 4)(*[_p(self.disp[s]) for s in range(S)])
            self._color_ptrs = (ctypes.c_void_p * 4)(*[_p(self.color[s]) for s in range(S)])
            plan.keep += [self._disp_ptrs, self._color_ptrs]
            plan.simple('smooth_masked', lib.clslam_smooth_masked, ctypes.cast(self._disp_ptrs, ctypes.c_void_p),
                        ctypes.cast(self._color_ptrs, ctypes.c_void_p), _p(self.mask_pos), _p(self.sample_w), b, h, w, S,
                        float(disparity_smoothness), _p(self.losses), _p(self.sm_idx), _p(self.sm_val), _p(self.sm_uniform))
        self.n_fwd_ops = len(plan.ops)

        # =================================== backward ====================================================================
        plan.add('photo_bwd', lambda stream, ws: lib.clslam_photo_bwd(ctypes.byref(pa), stream))
        self.dz_disp = {}
        for s in range(S):
            self.dz_disp[s] = plan.alloc(b, 1, h >> s, w >> s)
            plan.simple(f'disp_grad_{s}', lib.clslam_disp_grad, _p(self.ddisp_full[s]), _p(self.disp[s]), _p(self.smooth_grad), batch_global,
                        1 if (has_sample0 and not mask_dynamic) else 0, b, h, w, s, _p(self.dz_disp[s]))
            if mask_dynamic:
                plan.simple(f'smooth_masked_apply_{s}', lib.clslam_smooth_masked_apply, _p(self.disp[s]), _p(self.sm_idx), _p(self.sm_val),
                            _p(self.sm_uniform), b, h, w, s, _p(self.dz_disp[s]))
        self.dz = {k: plan.alloc(*v.shape) for k, v in self.y.items()}
        dfeat: Dict[int, torch.Tensor] = {}          # d(loss) / d(encoder feature i) from the decoder
        for i in range(0, 5):
            y1 = self.y[(i, 1)]
            hi_, wi_ = y1.shape[1], y1.shape[2]
            if i < S:
                keyd = (dd, f'dispconv_{i}.conv')
                plan.need_ws(lib.clslam_dispconv_wgrad_workspace_bytes(b, hi_, wi_, DEC_CH[i]))
                plan.with_ws(f'dispconv_{i}.wgrad', lib.clslam_dispconv_wgrad, _p(y1), _p(self.dz_disp[i]), flat.wgrad_ptr(keyd),
                             flat.bgrad_ptr(keyd), b, hi_, wi_, DEC_CH[i])
            if i == 0:
                plan.simple('dispconv_0.dgrad', lib.clslam_dispconv_dgrad_act, _p(self.dz_disp[0]), flat.weight_ptr((dd, 'dispconv_0.conv')),
                            _p(y1), L.ACT_ELU, _p(self.dz[(0, 1)]), b, hi_, wi_, DEC_CH[0])
            d1, y0, skip = self.dec_desc[(i, 1)]
            key1 = (dd, f'upconv_{i}_1.conv.conv')
            plan.conv_wgrad(f'upconv_{i}_1.wgrad', d1, y0, skip, self.dz[(i, 1)], flat.wgrad_ptr(key1), flat.bgrad_ptr(key1))
            plan.conv_dgrad(f'upconv_{i}_1.dgrad', d1, self.dz[(i, 1)], flat.weight_ptr(key1), y0, L.ACT_ELU, None, None, self.dz[(i, 0)])
            if skip is not None:          # the part of the input gradient the adaptation step never needs: into the encoder feature
                dfeat[i - 1] = plan.alloc(*skip.shape)
                self._dgrad_ex(f'upconv_{i}_1.dgrad_skip', d1, self.dz[(i, 1)], flat.weight_ptr(key1), DEC_CH[i], skip.shape[3], dfeat[i - 1])
            d0, xin = self.dec_desc[(i, 0)]
            key0 = (dd, f'upconv_{i}_0.conv.conv')
            plan.conv_wgrad(f'upconv_{i}_0.wgrad', d0, xin, None, self.dz[(i, 0)], flat.wgrad_ptr(key0), flat.bgrad_ptr(key0))
            if i < 4:
                has_disp = (i + 1) < S
                plan.conv_dgrad(f'upconv_{i}_0.dgrad', d0, self.dz[(i, 0)], flat.weight_ptr(key0), self.y[(i + 1, 1)], L.ACT_ELU,
                                self.dz_disp[i + 1] if has_disp else None,
                                flat.weight_ptr((dd, f'dispconv_{i + 1}.conv')) if has_disp else None, self.dz[(i + 1, 1)])
            else:
                dfeat[4] = plan.alloc(*feats[4].shape)
                self._dgrad_ex('upconv_4_0.dgrad_feat', d0, self.dz[(4, 0)], flat.weight_ptr(key0), 0, ENC_CH[4], dfeat[4])
        # pose decoder
        self.dz_pose = {name: plan.alloc(*self.pose_act[name].shape) for name in ('squeeze', 'pose_0', 'pose_1', 'pose_2')}
        plan.simple('pose_head_bwd', lib.clslam_pose_head_bwd, _p(self.dproj_partials), S * self.n_tiles, _p(self.k_mat), _p(self.axis_angle),
                    _p(self.translation), _p(self.rel_dist), _p(self.sample_w), self.vel_scale, b, hp, wp, self.invert_mask,
                    _p(self.dz_pose['pose_2']))
        prev = {'pose_2': 'pose_1', 'pose_1': 'pose_0', 'pose_0': 'squeeze'}
        for name in ('pose_2', 'pose_1', 'pose_0', 'squeeze'):
            d, xin = self.pose_desc[name]
            plan.conv_wgrad(f'{name}.wgrad', d, xin, None, self.dz_pose[name], flat.wgrad_ptr((pd, name)), flat.bgrad_ptr((pd, name)))
            if name in prev:
                plan.conv_dgrad(f'{name}.dgrad', d, self.dz_pose[name], flat.weight_ptr((pd, name)), self.pose_act[prev[name]], L.ACT_RELU,
                                None, None, self.dz_pose[prev[name]])
        dpf4 = plan.alloc(*pf4.shape)
        d_sq, _ = self.pose_desc['squeeze']
        self._dgrad_ex('squeeze.dgrad_feat', d_sq, self.dz_pose['squeeze'], flat.weight_ptr((pd, 'squeeze')), 0, ENC_CH[4], dpf4)
        # encoders
        self._encoder_backward('depth_enc', self.denc, depth_flat, dfeat, b, 1)
        self._encoder_backward('pose_enc', self.penc, pose_flat, {4: dpf4}, 2 * b, 2)
        self.n_bwd_ops = len(plan.ops)
        plan.finalize()
        self._masks: Dict[str, Optional[torch.Tensor]] = {}
        # nn.BatchNorm2d counts its training-mode calls (state_dict entry `num_batches_tracked`): one per depth-encoder module,
        # two per pose-encoder module and step (two pose passes, depth_pose_prediction.py:956-963)
        self._bn_counters = [(k, [m.num_batches_tracked for m in enc.modules() if hasattr(m, 'num_batches_tracked')])
                             for k, enc in ((1, depth_enc), (2, pose_enc))]

    # ----------------------------------------------------------------------------------------------------------------------
    def out_view(self, key, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, nel, shape = self.out_views[key]
        return (self.out_arena if arena is None else arena)[off:off + nel].view(*shape)

    def _dgrad_ex(self, name, d, dz, w_ptr, ci_off, ci_n, dx, y_prev=None, act_prev=L.ACT_NONE) -> None:
        plan = self.plan
        plan.need_ws(self.lib.clslam_conv_dgrad_ex_workspace_bytes(ctypes.byref(d), ci_n))
        plan.add(name, lambda stream, ws, *a: self.lib.clslam_conv_dgrad_ex(ctypes.byref(a[0]), a[1], a[2], a[3], a[4], a[5], a[6], a[7], ws,
                                                                             plan.ws_bytes, stream),
                 d, _p(dz), w_ptr, ci_off, ci_n, _p(y_prev), act_prev, _p(dx))

    def _bn_fwd(self, name, ef: EncoderFlat, key, mod, z, residual, relu, n, groups, y, saved) -> None:
        plan, lib = self.plan, self.lib
        c = z.shape[3]
        hw = z.shape[1] * z.shape[2]
        gamma, beta, _, _ = ef.bn_ptrs(key)
        mean, invstd = plan.alloc(groups, c), plan.alloc(groups, c)
        plan.need_ws(lib.clslam_bn_workspace_bytes(groups, c))
        plan.add(name, lambda stream, ws, *a: lib.clslam_bn_fwd(*a, ws, plan.ws_bytes, stream),
                 _p(z), gamma, beta, mod.running_mean.data_ptr(), mod.running_var.data_ptr(), 1 if self.bn_training else 0, BN_MOMENTUM, BN_EPS,
                 _p(residual), 1 if relu else 0, n, hw, c, groups, _p(y), _p(mean), _p(invstd))
        saved.update(mean=mean, invstd=invstd)

    def _bn_bwd(self, name, ef: EncoderFlat, key, rec, dy, relu, n, groups, dx, g_out=None) -> None:
        plan, lib = self.plan, self.lib
        z = rec['z']
        c, hw = z.shape[3], z.shape[1] * z.shape[2]
        gamma, _, dgamma, dbeta = ef.bn_ptrs(key)
        plan.need_ws(lib.clslam_bn_workspace_bytes(groups, c))
        plan.add(name, lambda stream, ws, *a: lib.clslam_bn_bwd(*a, ws, plan.ws_bytes, stream),
                 _p(z), _p(dy), _p(rec['y']), 1 if relu else 0, gamma, _p(rec['mean']), _p(rec['invstd']), 1 if self.bn_training else 0, n, hw,
                 c, groups, _p(dx), dgamma, dbeta, _p(g_out))

    def _encoder_forward(self, tag, enc: ResnetEncoder, ef: EncoderFlat, x_staged, n, groups):
        """ResNet-18 trunk with explicit BatchNorm (resnet_encoder.py:118-124, torchvision BasicBlock); every intermediate the
        backward pass needs is kept"""
        plan, lib = self.plan, self.lib
        r = enc.resnet
        h, w = x_staged.shape[1], x_staged.shape[2]
        rec = {'blocks': [], 'feats': []}

        def conv_bn(name, key_conv, key_bn, mod_bn, x, ci, co, k, stride, pad, residual, relu):
            hi, wi = x.shape[1], x.shape[2]
            d = plan.desc(n, hi, wi, ci, co, k, stride, pad, act=L.ACT_NONE)
            z = plan.alloc(n, d.ho, d.wo, co)
            plan.conv_fwd(f'{tag}.{name}', d, x, None, ef.w_ptr(key_conv), None, None, None, z)
            y = plan.alloc(n, d.ho, d.wo, co)
            saved = {'d': d, 'x': x, 'z': z, 'y': y, 'conv': key_conv, 'bn': key_bn}
            self._bn_fwd(f'{tag}.{name}.bn', ef, key_bn, mod_bn, z, residual, relu, n, groups, y, saved)
            return saved

        c1 = conv_bn('conv1', 'conv1', 'bn1', r.bn1, x_staged, ef.cin_pad, 64, 7, 2, 3, None, True)
        rec['conv1'] = c1
        f0 = c1['y']
        rec['feats'].append(f0)
        hp, wp = conv_out(f0.shape[1], 3, 2, 1), conv_out(f0.shape[2], 3, 2, 1)
        pooled = plan.alloc(n, hp, wp, 64)
        plan.simple(f'{tag}.maxpool', lib.clslam_maxpool3x3s2, _p(f0), _p(pooled), n, f0.shape[1], f0.shape[2], 64)
        rec['pooled'] = pooled
        x, cin = pooled, 64
        for li, planes in enumerate((64, 128, 256, 512), start=1):
            for bi, blk in enumerate(getattr(r, f'layer{li}')):
                p = f'layer{li}.{bi}'
                stride = 2 if (li > 1 and bi == 0) else 1
                a = conv_bn(f'{p}.conv1', f'{p}.conv1', f'{p}.bn1', blk.bn1, x, cin, planes, 3, stride, 1, None, True)
                if blk.downsample is not None:
                    ds = conv_bn(f'{p}.downsample', f'{p}.ds', f'{p}.dsbn', blk.downsample[1], x, cin, planes, 1, stride, 0, None, False)
                    idt = ds['y']
                else:
                    ds, idt = None, x
                bb = conv_bn(f'{p}.conv2', f'{p}.conv2', f'{p}.bn2', blk.bn2, a['y'], planes, planes, 3, 1, 1, idt, True)
                rec['blocks'].append({'name': p, 'x': x, 'a': a, 'b': bb, 'ds': ds, 'li': li, 'bi': bi})
                x, cin = bb['y'], planes
            rec['feats'].append(x)
        return rec

    def _encoder_backward(self, tag, rec, ef: EncoderFlat, dfeat: Dict[int, torch.Tensor], n, groups) -> None:
        """backward of _encoder_forward.  dfeat[i]: gradient that reaches encoder feature i from outside the encoder (decoder skip
        connections, the pose decoder); features without an entry only receive gradient from the following block"""
        plan, lib = self.plan, self.lib

        def conv_back(name, s, dz, need_dx):
            d, x = s['d'], s['x']
            plan.conv_wgrad(f'{tag}.{name}.wgrad', d, x, None, dz, ef.dw_ptr(s['conv']), None)
            if not need_dx:
                return None
            dx = plan.alloc(*x.shape)
            self._dgrad_ex(f'{tag}.{name}.dgrad', d, dz, ef.w_ptr(s['conv']), 0, d.ci, dx)
            return dx

        def summed(name, parts, like):
            parts = [t for t in parts if t is not None]
            if len(parts) == 1:
                return parts[0]
            out = plan.alloc(*like.shape)
            plan.simple(f'{tag}.{name}.sum', lib.clslam_add3, _p(parts[0]), _p(parts[1]), _p(parts[2]) if len(parts) > 2 else None, _p(out),
                        out.numel())
            return out

        dy = dfeat.get(4)          # gradient wrt the last block's output (feature 4), before its ReLU mask
        for blk in reversed(rec['blocks']):
            p, a, bb, ds = blk['name'], blk['a'], blk['b'], blk['ds']
            # conv2 + bn2 (+ identity) + relu
            dz2 = plan.alloc(*bb['z'].shape)
            g = plan.alloc(*bb['z'].shape)          # dy * relu'(y): also the gradient of the identity branch
            self._bn_bwd(f'{tag}.{p}.bn2.bwd', ef, bb['bn'], bb, dy, True, n, groups, dz2, g_out=g)
            dmid = conv_back(f'{p}.conv2', bb, dz2, True)
            dz1 = plan.alloc(*a['z'].shape)
            self._bn_bwd(f'{tag}.{p}.bn1.bwd', ef, a['bn'], a, dmid, True, n, groups, dz1)
            dx1 = conv_back(f'{p}.conv1', a, dz1, True)
            if ds is not None:
                dzd = plan.alloc(*ds['z'].shape)
                self._bn_bwd(f'{tag}.{p}.dsbn.bwd', ef, ds['bn'], ds, g, False, n, groups, dzd)
                dxi = conv_back(f'{p}.downsample', ds, dzd, True)
            else:
                dxi = g
            # the block input is encoder feature li - 1 when this is the first block of layer li (li >= 2)
            outside = dfeat.get(blk['li'] - 1) if (blk['bi'] == 0 and blk['li'] >= 2) else None
            dy = summed(p, [dx1, dxi, outside], blk['x'])
        # layer1.0's input is the max-pooled feature 0
        c1 = rec['conv1']
        f0 = c1['y']
        df0 = plan.alloc(*f0.shape)
        plan.simple(f'{tag}.maxpool.bwd', lib.clslam_maxpool3x3s2_bwd, _p(f0), _p(dy), _p(df0), n, f0.shape[1], f0.shape[2], 64)
        df0 = summed('feat0', [df0, dfeat.get(0)], f0)
        dz_c1 = plan.alloc(*c1['z'].shape)
        self._bn_bwd(f'{tag}.bn1.bwd', ef, 'bn1', c1, df0, True, n, groups, dz_c1)
        conv_back('conv1', c1, dz_c1, False)

    # ---- execution -------------------------------------------------------------------------------------------------------
    def _count_bn_calls(self) -> None:
        if self.bn_training:
            for k, counters in self._bn_counters:
                torch._foreach_add_(counters, k)

    def forward(self) -> None:
        self.plan.run(0, self.n_fwd_ops)
        self._count_bn_calls()

    def backward(self) -> None:
        self.plan.run(self.n_fwd_ops, self.n_bwd_ops)

    def forward_backward(self, timed: bool = False) -> None:
        self.plan.run(0, self.n_bwd_ops, timed=timed)
        self._count_bn_calls()

    def adam_hparams(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
        self.adam_lr = float(lr)

    def mask_frozen(self) -> None:
        """zero the gradients of parameters with requires_grad = False (the reference's optimiser skips them: the 'bn'-named
        BatchNorm parameters after _set_adapt(freeze_encoder=False), :813-816)"""
        for tag, ef in (('depth', self.depth_flat), ('pose', self.pose_flat)):
            m = self._masks.get(tag)
            if m is not None:
                ef.grads.mul_(m)

    def set_frozen_masks(self) -> None:
        for tag, ef in (('depth', self.depth_flat), ('pose', self.pose_flat)):
            m = ef.frozen_mask()
            self._masks[tag] = None if bool((m == 1).all()) else m

    def adam(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
        """Adam over the decoders' flat buffer and both encoders' flat buffers; ONE step counter (torch.optim.Adam keeps a
        counter per tensor, all equal)"""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.mask_frozen()
        f = self.flat
        bufs = [(f.params, f.grads, f.exp_avg, f.exp_avg_sq, f.numel)]
        for ef in (self.depth_flat, self.pose_flat):
            bufs.append((ef.params, ef.grads, ef.exp_avg, ef.exp_avg_sq, ef.numel))
        for p, g, m, v, n in bufs:
            L.check(self.lib.clslam_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr, beta1, beta2, eps,
                                              L.ADAM_NO_TICK, f.step_dev.data_ptr(), stream), 'adam')
        L.check(self.lib.clslam_adam_tick(f.step_dev.data_ptr(), stream), 'adam_tick')
