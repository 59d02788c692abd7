"""Drop-in ``DepthPosePrediction`` for CL-SLAM's online adaptation, running on hand-written
sm_100a CUDA through the C ABI in ``include/clslam_b200.h``.

Mirrors the surface of ``depth_pose_prediction/depth_pose_prediction.py:38`` that
``slam/slam.py`` uses: constructor ``(dataset_config, config, use_online)``,
``adapt`` (:291), ``_process_batch`` (:906), ``predict_pose`` (:628), ``load_model`` (:705),
``save_model`` (:669), ``_set_eval`` / ``_set_adapt`` (:802, :807), attributes ``models``,
``device``, ``min_depth``, ``max_depth``, the batch-dict schema and the 19 loss keys
(SURVEY.md App. C).  There is no CPU / PyTorch fallback: compute methods raise when CUDA or the
extension is missing.
"""
from __future__ import annotations

import os
import shutil
import warnings
from pathlib import Path
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn, optim

from . import _lib as L
from .engine import H2D_OVERLAP, AdaptEngine, DepthRunner, EncoderRunner, Plan, PreparedEncoder, _p
from .networks import DepthDecoder, EncoderFlat, FlatParams, PoseDecoder, ResnetEncoder
from .train_engine import TrainEngine
from .sharding import global_sample_weights, shard_range

# data-parallel: all-reduce the gradients in buckets inside the step graph, overlapped with the backward pass (default);
# CLSLAM_GRAD_BUCKETS=0 restores round 1's graph -> eager all-reduce of the whole buffer -> graph
GRAD_BUCKETS = int(os.environ.get('CLSLAM_GRAD_BUCKETS', '1'))        # 0 / 1 / 2 / 5: see AdaptEngine (measured at 2 GPUs, profiles/r02_dp_buckets.md: 1 is fastest)
LANE0_PRIORITY = int(os.environ.get('CLSLAM_LANE0_PRIORITY', '0'))
BUCKET_ADAM = os.environ.get('CLSLAM_BUCKET_ADAM', '0') == '1'

LOSS_SLOTS = {'reprojection_loss': 0, 'smooth_loss': 4, 'reg_loss': 8, 'depth_loss': 12}


class FlatAdam:
    """torch.optim.Adam-shaped facade (state_dict format, param_groups, LR scheduler target) over
    the fused flat-buffer Adam kernel.  The wrapped torch optimiser is never stepped; it only
    carries hyper-parameters and the (de)serialisation format of depth_pose_prediction.py:691-697."""

    def __init__(self, all_params: List[nn.Parameter], lr: float, flat: FlatParams, trainable: List[nn.Parameter],
                 encoder_flats: Optional[List[EncoderFlat]] = None):
        self.torch_optimizer = optim.Adam(all_params, lr)
        self.flat = flat
        self.trainable = trainable      # the 36 decoder tensors, in flat-buffer order (weight, bias, ...)
        self.encoder_flats = encoder_flats or []          # trained by train() / _set_adapt(freeze_encoder=False) only
        self.step_fn: Optional[Callable[[], None]] = None
        self._spans = []                # (param, offset, numel, is_weight, info)
        for (net, path), info in flat.index.items():
            self._spans.append((info, True))
            self._spans.append((info, False))

    @property
    def param_groups(self):
        return self.torch_optimizer.param_groups

    def _state_views(self, buf: Tensor):
        out = []
        for info, is_w in self._spans:
            if is_w:
                out.append(FlatParams.view_weight(buf, info))
            else:
                out.append(buf[info['b_off']:info['b_off'] + info['co']])
        return out

    def _triples(self):
        out = list(zip(self.trainable, self._state_views(self.flat.exp_avg), self._state_views(self.flat.exp_avg_sq)))
        for ef in self.encoder_flats:          # only once the encoders have been trained (their moment buffers exist)
            out += [(p, m, v) for p, m, v in ef.state_triples() if p.requires_grad]
        return out

    def state_dict(self):
        steps = int(self.flat.step_dev.item())
        st = self.torch_optimizer.state
        if steps > 0:
            for p, m, v in self._triples():
                st[p] = {'step': torch.tensor(float(steps)), 'exp_avg': m.detach().clone().contiguous(),
                         'exp_avg_sq': v.detach().clone().contiguous()}
        return self.torch_optimizer.state_dict()

    def load_state_dict(self, sd) -> None:
        self.torch_optimizer.load_state_dict(sd)
        st = self.torch_optimizer.state
        if any(p in st for ef in self.encoder_flats for p in ef.enc.parameters()):
            for ef in self.encoder_flats:      # a checkpoint of a trained encoder: its moments need somewhere to live
                ef.ensure_training_buffers()
        steps = 0
        with torch.no_grad():
            for p, m, v in self._triples():
                if p in st and 'exp_avg' in st[p]:
                    m.copy_(st[p]['exp_avg']); v.copy_(st[p]['exp_avg_sq'])
                    steps = max(steps, int(float(st[p]['step'])))
            self.flat.step_dev.fill_(steps)

    def step(self) -> None:
        """``optimizer.step()`` of the reference's training loop (:256): Adam on the gradients of the last ``backward()``"""
        if self.step_fn is None:
            raise RuntimeError('FlatAdam.step(): no backward pass has been run')
        self.step_fn()

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.flat.grads.zero_()
        for ef in self.encoder_flats:
            if ef.grads is not None:
                ef.grads.zero_()


# inputs that only the photometric chain / the losses read (H2D_OVERLAP: they travel on the copy stream while the networks run)
LATE_INPUT_KEYS = frozenset([('rgb', 0, 0), ('rgb', -1, 0), ('rgb', 1, 0), ('rgb', 0, 1), ('rgb', 0, 2), ('rgb', 0, 3),
                             ('relative_distance', 0), ('relative_distance', 1)])


class _NetSet:
    """One copy of the four networks with everything derived from it: the flat trainable buffers, the kernel-layout copies
    of the frozen encoders, the launch plans and their CUDA graphs.  The predictor has one (``models``) or, with
    ``use_online=True``, two (``online_models``, depth_pose_prediction.py:139-152)."""

    def __init__(self, owner: 'DepthPosePrediction', tag: str):
        dev = owner.device
        self.tag = tag
        self.models: Dict[str, nn.Module] = {}
        self.models['depth_encoder'] = ResnetEncoder(owner.resnet_depth, owner.resnet_pretrained, device=dev)
        self.models['depth_decoder'] = DepthDecoder(self.models['depth_encoder'].num_ch_encoder, owner.scales)
        self.models['pose_encoder'] = ResnetEncoder(owner.resnet_pose, owner.resnet_pretrained, owner.num_pose_frames, device=dev)
        self.models['pose_decoder'] = PoseDecoder()
        self.flat = FlatParams(self.models['depth_decoder'], self.models['pose_decoder'], dev)
        # encoder weights / BatchNorm affine parameters as views of flat buffers too (the full training path updates them in place)
        self.enc_flat = {n: EncoderFlat(self.models[n], dev) for n in ('depth_encoder', 'pose_encoder')}
        self.models['depth_encoder'].bind(lambda x: owner._run_encoder(self, 'depth_encoder', x))
        self.models['pose_encoder'].bind(lambda x: owner._run_encoder(self, 'pose_encoder', x))
        self.parameters: List[nn.Parameter] = []
        for m in self.models.values():
            self.parameters += list(m.parameters())
        trainable = list(self.models['depth_decoder'].parameters()) + list(self.models['pose_decoder'].parameters())
        self.optimizer = FlatAdam(self.parameters, owner.learning_rate, self.flat, trainable, list(self.enc_flat.values()))
        self.train_engines: Dict[Tuple[int, bool], TrainEngine] = {}
        self.prep: Dict[str, PreparedEncoder] = {}
        self.engines: Dict[int, AdaptEngine] = {}
        self.enc_runners: Dict[Tuple[str, int, int, int], EncoderRunner] = {}
        self.pose_runner = None
        self.depth_runners: Dict[Tuple[int, int, int], DepthRunner] = {}
        self.graphs: Dict[Tuple, Any] = {}
        self.steps_run: Dict[int, int] = {}
        self.weights_key: Dict[int, Any] = {}
        self.adapt_flags_set = False          # False | "bn_only" | "frozen" (requires_grad is only ever switched off)
        self.enc_trainable = True

    def drop_plans(self) -> None:
        """plans hold raw pointers into the prepared encoder tensors: forget them (and everything captured from them)"""
        self.engines.clear(); self.enc_runners.clear(); self.graphs.clear(); self.depth_runners.clear()
        # (the training engines read the encoders' flat buffers and the modules' running statistics, whose addresses never
        # change: they survive; their graphs went with self.graphs)
        self.pose_runner = None
        self.weights_key.clear()
        self.steps_run.clear()          # a fresh engine runs its first step eagerly again (loads kernels, builds tensor maps)


class DepthPosePrediction:
    def __init__(self, dataset_config, config, use_online: bool = False, *, data_parallel: bool = False,
                 use_cuda_graph: bool = True, emit_outputs: bool = True, copy_outputs: bool = True,
                 device: Optional[torch.device] = None):
        # Parameters (same names as the reference, :41-68) --------------------------------------
        self.config_file = config.config_file
        self.dataset_type = dataset_config.dataset
        self.dataset_path = dataset_config.dataset_path
        self.height = dataset_config.height
        self.width = dataset_config.width
        self.train_set = config.train_set
        self.val_set = config.val_set
        self.resnet_depth = config.resnet_depth
        self.resnet_pose = config.resnet_pose
        self.resnet_pretrained = config.resnet_pretrained
        self.scales = tuple(config.scales)
        self.learning_rate = config.learning_rate
        self.scheduler_step_size = config.scheduler_step_size
        self.batch_size = config.batch_size
        self.num_workers = config.num_workers
        self.num_epochs = config.num_epochs
        self.min_depth = config.min_depth
        self.max_depth = config.max_depth
        self.disparity_smoothness = config.disparity_smoothness
        self.velocity_loss_scaling = config.velocity_loss_scaling
        self.mask_dynamic = config.mask_dynamic
        self.log_path = config.log_path
        self.save_frequency = config.save_frequency
        self.save_val_depth = config.save_val_depth
        self.save_val_depth_batches = config.save_val_depth_batches
        self.multiple_gpus = config.multiple_gpus
        self.gpu_ids = config.gpu_ids
        self.load_weights_folder = config.load_weights_folder
        self.use_wandb = False
        self.is_trained = False
        self.frame_ids = (0, -1, 1)          # :75
        self.num_pose_frames = 2
        self.epoch = 0

        self._validate_sets(dataset_config)
        if self.multiple_gpus:
            raise ValueError('nn.DataParallel (multiple_gpus) is replaced by data_parallel=True: one process per '
                             'GPU under torch.distributed (SURVEY.md 8e)')
        if self.load_weights_folder is not None:
            self.load_weights_folder = Path(self.load_weights_folder).absolute()
        self.use_online = use_online

        # Device / process group ------------------------------------------------------------------
        self.rank, self.world = 0, 1
        if data_parallel:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError('data_parallel=True needs an initialised torch.distributed process group')
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.data_parallel = data_parallel
        if device is not None:
            self.device = torch.device(device)
        elif torch.cuda.is_available():
            if self.gpu_ids is not None and not data_parallel:
                torch.cuda.set_device(self.gpu_ids[0])          # "the main GPU" of :122-124
            self.device = torch.device('cuda', torch.cuda.current_device())
            if self.gpu_ids is None:
                self.gpu_ids = (self.device.index,)
        else:
            self.device = torch.device('cpu')      # parameter containers / checkpoints only; compute raises
        self.use_cuda_graph = use_cuda_graph
        self.emit_outputs = emit_outputs
        # The reference returns fresh tensors from every call (:906-923).  The engine writes its outputs into static buffers,
        # so the public methods hand out ONE clone of the output arena per call (a single D2D copy).  copy_outputs=False
        # returns views of the static buffers instead: valid until the next adapt / predict call on this object.
        self.copy_outputs = copy_outputs
        # Slam.step converts every one of the 19 losses with `float(v.squeeze().cpu()...)` (slam/slam.py:186-188): 19 device
        # syncs per frame.  The class already reads the whole loss vector back once per call for the NaN check (:1115); with
        # losses_on_host=True the returned loss tensors are views of THAT host copy, so the caller's `.cpu()` calls are free.
        # Off by default: the reference returns device tensors.
        self.losses_on_host = False

        # Networks (parameter containers; arithmetic lives in the CUDA kernels) --------------------
        self._main = _NetSet(self, 'main')
        self.models: Dict[str, nn.Module] = self._main.models
        self.parameters_to_train: List[nn.Parameter] = self._main.parameters
        self.optimizer = self._main.optimizer
        self.lr_scheduler = optim.lr_scheduler.StepLR(self.optimizer.torch_optimizer, self.scheduler_step_size, 0.1)
        # second copy of the four networks (:139-152, :206-207): evaluated by _process_batch / predict_pose(use_online=True)
        self._online = _NetSet(self, 'online') if use_online else None
        self.online_models: Dict[str, nn.Module] = self._online.models if use_online else {}
        self.online_parameters_to_train: List[nn.Parameter] = self._online.parameters if use_online else []
        self.online_optimizer = self._online.optimizer if use_online else None
        self.train_loader, self.val_loader = None, None

        # Runtime state ---------------------------------------------------------------------------
        self._aux_dev: Dict[Any, Tensor] = {}
        self._deferred: List[Tuple[Tensor, Tensor]] = []
        self._copy_stream = None
        self.noise_provider: Optional[Callable[[int, Tuple[int, ...]], Tensor]] = None
        self.stage_all_inputs = True          # move every dict value to the device like :916-917
        self._mode = 'eval'
        self._late_event = None

    # the main network set under its round-1 names (tests, bench.py)
    @property
    def _flat(self) -> FlatParams:
        return self._main.flat

    @property
    def _engines(self) -> Dict[int, AdaptEngine]:
        return self._main.engines

    def close(self) -> None:
        """release the launch plans and CUDA graphs.  Data-parallel runs must call this before
        ``torch.distributed.destroy_process_group()``: the step graph holds captured NCCL kernels, and tearing the communicator
        down underneath a live graph hangs."""
        for ns in self._sets():
            for g in ns.graphs.values():
                try:
                    g.reset()
                except Exception:   # noqa: BLE001
                    pass
            ns.drop_plans()
            ns.train_engines.clear()
        self._late_event = None
        self._last_engine = None
        if self.device.type == 'cuda':
            torch.cuda.synchronize(self.device)

    def _sets(self) -> List[_NetSet]:
        return [self._main] + ([self._online] if self._online is not None else [])

    def _set_of(self, use_online: bool) -> _NetSet:
        if use_online:
            if self._online is None:
                raise KeyError('depth_encoder')      # the reference indexes its empty online_models dict (:931)
            return self._online
        return self._main

    # ------------------------------------------------------------------------------------------
    def _validate_sets(self, dataset_config) -> None:
        """train_set / val_set normalisation and checks of :85-110 (ValueError on bad config)."""
        if isinstance(self.train_set, list):
            self.train_set = tuple(self.train_set)
        if isinstance(self.val_set, list):
            self.val_set = tuple(self.val_set)
        kind = dataset_config.dataset
        if kind == 'Kitti':
            if isinstance(self.val_set, int):
                self.val_set = (self.val_set,)
            if isinstance(self.train_set, str):
                if self.train_set != 'all':
                    raise ValueError('train_set of KITTI only accepts these strings: ["all"]')
                self.train_set = tuple(s for s in range(11) if s not in self.val_set and s != 3)
            elif isinstance(self.train_set, int):
                self.train_set = (self.train_set,)
            want = int
        elif kind in ('Cityscapes', 'RobotCar'):
            if isinstance(self.train_set, str):
                self.train_set = (self.train_set,)
            if isinstance(self.val_set, str):
                self.val_set = (self.val_set,)
            want = str
        else:
            return
        for name, value in (('train_set', self.train_set), ('val_set', self.val_set)):
            if not (isinstance(value, tuple) and len(value) > 0 and isinstance(value[0], want)):
                raise ValueError(f'Passed invalid value for {name}')
        if self.gpu_ids is not None and len(self.gpu_ids) > 1:
            raise ValueError('Passed multiple GPU IDs without activating multi-GPU support.')

    def _require_cuda(self) -> None:
        if self.device.type != 'cuda':
            raise RuntimeError('clslam_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        L.load()

    # ------------------------------------------------------------------------------------------
    # modes (:797-827)
    def _set_train(self) -> None:
        """:797-800: every network in training mode -- BatchNorm uses batch statistics and updates its running statistics;
        whatever has requires_grad is trained (the full training plan, clslam_b200/train_engine.py)"""
        for m in self.models.values():
            m.train()
        self._mode = 'train'

    # Both mode switches are called once per frame by Slam.step / adapt (slam/slam.py:143, :291-319).  Walking the four module
    # trees costs 0.4-0.6 ms of host time per call, in front of the step's H2D copies and launches, and is idempotent: the
    # modules only hold the parameters here (the CUDA path always evaluates BatchNorm with running statistics and never
    # differentiates the encoders), so the walk is skipped when nothing changed since the last one.
    def _all_eval(self) -> bool:
        return not any(m.training for m in self.models.values())

    def _set_eval(self) -> None:
        if not self._all_eval():
            for m in self.models.values():      # :802-805 (the online copy is not touched by _set_eval in the reference either)
                m.eval()
        self._mode = 'eval'

    def _set_adapt(self, freeze_encoder: bool = True) -> None:
        """:807-827.  BatchNorm in evaluation mode; 'bn'-named parameters frozen; with ``freeze_encoder`` every encoder
        parameter frozen too.  requires_grad is only ever switched OFF (the reference never re-enables it).  With trainable
        encoder convolutions left (freeze_encoder=False on a fresh object) the step runs on the full training plan."""
        for ns in self._sets():                 # :813-827: both copies
            if any(m.training for m in ns.models.values()):
                for m in ns.models.values():
                    m.eval()
            key = 'frozen' if freeze_encoder else 'bn_only'
            if ns.adapt_flags_set != 'frozen' and ns.adapt_flags_set != key:
                for model_name, model in ns.models.items():
                    for name, param in model.named_parameters():
                        if name.find('bn') != -1 or (freeze_encoder and 'encoder' in model_name):
                            param.requires_grad = False
                ns.adapt_flags_set = key
                ns.enc_trainable = any(p.requires_grad for n in ('depth_encoder', 'pose_encoder')
                                       for p in ns.models[n].resnet.parameters() if p.dim() == 4)
        self._mode = 'adapt_enc' if self._main.enc_trainable else 'adapt'

    # ------------------------------------------------------------------------------------------
    # engines
    def _prepared(self, net: str, ns: Optional[_NetSet] = None) -> PreparedEncoder:
        ns = ns or self._main
        enc = ns.models[net]
        prep = ns.prep.get(net)
        if prep is None or prep.version != enc.version:
            prep = PreparedEncoder(enc, self.device)
            ns.prep[net] = prep
            ns.drop_plans()
            self._late_event = None
        return prep

    def refresh_encoders(self) -> None:
        """Call after modifying encoder weights in place (load_state_dict does it automatically)."""
        for ns in self._sets():
            for net in ('depth_encoder', 'pose_encoder'):
                ns.models[net].version += 1

    def _engine(self, batch: int, ns: Optional[_NetSet] = None) -> AdaptEngine:
        self._require_cuda()
        ns = ns or self._main
        dp, pp = self._prepared('depth_encoder', ns), self._prepared('pose_encoder', ns)
        eng = ns.engines.get(batch)
        if eng is None and self.mask_dynamic:
            raise NotImplementedError('mask_dynamic=True (pre-training with dynamic-object masks, :1062-1068) runs on the training '
                                      'plan: _set_train() / train() / _set_adapt(freeze_encoder=False); the frozen-encoder '
                                      'adaptation and evaluation plans do not read masks')
        if eng is None:
            L.check(L.load().clslam_set_device(self.device.index or 0), 'set_device')
            if self.world > 1:
                b_global, has0 = batch * self.world, self.rank == 0
            else:
                b_global, has0 = batch, True
            eng = AdaptEngine(batch=batch, height=self.height, width=self.width, batch_global=b_global, has_sample0=has0,
                              depth_prep=dp, pose_prep=pp, flat=ns.flat, min_depth=self.min_depth,
                              max_depth=self.max_depth, disparity_smoothness=self.disparity_smoothness,
                              velocity_loss_scaling=self.velocity_loss_scaling, device=self.device,
                              emit_outputs=self.emit_outputs,
                              grad_allreduce=self._allreduce_tensor if (self.world > 1 and GRAD_BUCKETS > 0) else None,
                              grad_buckets=GRAD_BUCKETS, bucket_adam=BUCKET_ADAM)
            eng.net_set = ns
            eng.kind = 'adapt'
            ns.engines[batch] = eng
        return eng

    def _train_engine(self, batch: int, bn_training: bool, ns: Optional[_NetSet] = None) -> TrainEngine:
        """the full training plan (encoders trainable, BatchNorm explicit) for this local batch size"""
        self._require_cuda()
        ns = ns or self._main
        eng = ns.train_engines.get((batch, bn_training))
        if eng is None:
            L.check(L.load().clslam_set_device(self.device.index or 0), 'set_device')
            for ef in ns.enc_flat.values():
                ef.ensure_training_buffers()
            b_global, has0 = (batch * self.world, self.rank == 0) if self.world > 1 else (batch, True)
            eng = TrainEngine(batch=batch, height=self.height, width=self.width, batch_global=b_global, has_sample0=has0,
                              depth_enc=ns.models['depth_encoder'], pose_enc=ns.models['pose_encoder'],
                              depth_flat=ns.enc_flat['depth_encoder'], pose_flat=ns.enc_flat['pose_encoder'], flat=ns.flat,
                              min_depth=self.min_depth, max_depth=self.max_depth, disparity_smoothness=self.disparity_smoothness,
                              velocity_loss_scaling=self.velocity_loss_scaling, bn_training=bn_training, device=self.device,
                              emit_outputs=self.emit_outputs, mask_dynamic=self.mask_dynamic)
            eng.net_set = ns
            eng.kind = 'train' if bn_training else 'train_bn_eval'
            ns.train_engines[(batch, bn_training)] = eng
        eng.set_frozen_masks()          # requires_grad may have changed since the last call
        return eng

    def _run_encoder(self, ns: _NetSet, net: str, x: Tensor) -> List[Tensor]:
        """``models['depth_encoder'](image)`` (slam/slam.py:146): five NCHW-shaped feature maps."""
        self._require_cuda()
        prep = self._prepared(net, ns)
        n, _, h, w = x.shape
        key = (net, n, h, w)
        runner = ns.enc_runners.get(key)
        if runner is None:
            L.check(L.load().clslam_set_device(self.device.index or 0), 'set_device')
            runner = EncoderRunner(ns.models[net], prep, n, h, w, self.device, use_graph=self.use_cuda_graph)
            ns.enc_runners[key] = runner
        feats = runner(x.to(self.device, dtype=torch.float32))
        return [f.permute(0, 3, 1, 2).clone() for f in feats]

    # ------------------------------------------------------------------------------------------
    # input staging (:916-917) and weights
    def _static_map(self, eng: AdaptEngine) -> Dict[Any, Tensor]:
        m = {('rgb_aug', -1, 0): eng.rgb_aug[0], ('rgb_aug', 0, 0): eng.rgb_aug[1], ('rgb_aug', 1, 0): eng.rgb_aug[2],
             ('rgb', 0, 0): eng.target, ('rgb', -1, 0): eng.sources[0], ('rgb', 1, 0): eng.sources[1],
             ('camera_matrix', 0): eng.k_mat, ('inv_camera_matrix', 0): eng.inv_k,
             ('relative_distance', 0): eng.rel_dist[0], ('relative_distance', 1): eng.rel_dist[1]}
        for s in range(1, 4):
            m[('rgb', 0, s)] = eng.color[s]
        if getattr(eng, 'mask_dynamic', False):          # pre-training masks (:1066, :1081): training plan only
            for s in range(4):
                m[('mask', 0, s)] = eng.dyn_mask[s]
        return m

    def _stage_inputs(self, inputs: Dict[Any, Tensor], defer_unused: bool = False, ns: Optional[_NetSet] = None) -> AdaptEngine:
        """copies the dict to the device in place (:916-917).  With ``defer_unused`` the values the path does not read are
        only queued (``_flush_deferred`` moves them on a side stream after the step has been launched).

        The values the path reads land in the engine's static input buffers and the dict then holds views of those (in the
        caller's shape): they stay valid until the next call that stages a batch of the same size."""
        batch = inputs[('rgb_aug', 0, 0)].shape[0]
        if self._mode in ('train', 'adapt_enc') and (ns is None or ns is self._main):
            eng = self._train_engine(batch, self._mode == 'train', ns)
        else:
            eng = self._engine(batch, ns)
        smap = self._static_map(eng)
        self._deferred = []
        cuda = self.device.type == 'cuda'
        if defer_unused and cuda:
            self._stage_event = torch.cuda.Event()
            self._stage_event.record(torch.cuda.current_stream(self.device))
        # (data-parallel: only when the gradient exchange lives inside the plan, i.e. the step is graph-capturable as a whole)
        overlap = H2D_OVERLAP and defer_unused and cuda and (self.world == 1 or getattr(eng, 'grad_sync_in_plan', False))
        late: List[Tuple[Tensor, Tensor]] = []
        for key in list(inputs.keys()):
            val = inputs[key]
            dst = smap.get(key)
            if dst is not None:
                if not (val.is_cuda and val.data_ptr() == dst.data_ptr()):
                    src = val if val.shape == dst.shape else val.reshape(dst.shape)
                    if overlap and key in LATE_INPUT_KEYS:
                        late.append((dst, src))
                    else:
                        dst.copy_(src, non_blocking=True)
                inputs[key] = dst if (dst.shape == val.shape or dst.numel() != val.numel()) else dst.view(val.shape)
            elif self.stage_all_inputs and isinstance(val, Tensor):
                # keys the path does not read are still moved like :916-917, into cached device buffers (no allocation per call)
                buf = self._aux_dev.get(key)
                if buf is None or buf.shape != val.shape or buf.dtype != val.dtype:
                    buf = torch.empty(val.shape, dtype=val.dtype, device=self.device)
                    self._aux_dev[key] = buf
                if defer_unused:
                    self._deferred.append((buf, val))
                else:
                    buf.copy_(val, non_blocking=True)
                inputs[key] = buf
        missing = [k for k in smap if k not in inputs and not (k[0] == 'relative_distance' and eng.vel_scale == 0.0)]
        if missing:
            raise KeyError(f'batch dict lacks keys the adaptation path reads: {missing}')
        self._late_event = None
        if overlap and late:
            # the photometric inputs follow the network inputs on the copy stream; the second graph of the step waits for them
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            cur = torch.cuda.current_stream(self.device)
            self._copy_stream.wait_stream(cur)        # earlier readers of these buffers (the previous step) and the copies above
            with torch.cuda.stream(self._copy_stream):
                for dst, src in late:
                    dst.copy_(src, non_blocking=True)
            self._late_event = torch.cuda.Event()
            self._late_event.record(self._copy_stream)
        return eng

    def _flush_deferred(self) -> None:
        """H2D of the unused dict values on a side stream, overlapping the step; the caller's stream waits for them"""
        if not self._deferred:
            return
        if self.device.type != 'cuda':          # parameter-container mode (host-logic tests): nothing to overlap
            for buf, val in self._deferred:
                buf.copy_(val)
            self._deferred = []
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_event(self._stage_event)     # readers of the cached buffers queued before this call are done
        with torch.cuda.stream(self._copy_stream):
            for buf, val in self._deferred:
                buf.copy_(val, non_blocking=True)
        cur.wait_stream(self._copy_stream)
        self._deferred = []

    def _set_weights(self, eng: AdaptEngine, loss_sample_weights: Optional[Tensor]) -> None:
        b_global = eng.b_global
        # the weights rarely change between calls: upload only when they do (a pageable H2D copy stalls the host)
        key = None if loss_sample_weights is None else tuple(float(v) for v in loss_sample_weights.detach().flatten().tolist())
        wk = eng.net_set.weights_key
        slot = (eng.b, getattr(eng, 'kind', 'adapt'))          # the adaptation and the training engine have their own buffers
        if wk.get(slot, 'unset') == (key, self.batch_size):
            return
        wk[slot] = (key, self.batch_size)
        if loss_sample_weights is None:
            w = torch.ones(self.batch_size) / self.batch_size          # :1032 (uses config batch_size)
        else:
            w = loss_sample_weights.detach().float().cpu()
        if w.numel() != b_global:
            if b_global == 1:
                w = w.sum().reshape(1)          # (1,) * (batch_size,) broadcast then .sum() in the reference
            else:
                raise ValueError(f'{w.numel()} loss weights for a global batch of {b_global}')
        lo, hi = shard_range(b_global, self.rank, self.world)
        eng.sample_w_global.copy_(w, non_blocking=True)
        eng.sample_w.copy_(w[lo:hi], non_blocking=True)

    def _draw_noise(self, eng: AdaptEngine) -> None:
        """tie-break noise of :1055-1056: one draw per scale from the global torch RNG, in scale order"""
        if self.noise_provider is None:
            # one RNG launch over the scale-major buffer instead of four (same distribution, same number of values; the
            # reference draws on its own device generator, so the streams never matched bitwise -- parity runs inject the
            # noise through ``noise_provider``)
            eng.noise_all.normal_()
            return
        for s in range(4):
            n = self.noise_provider(s, (eng.b_global, 2, self.height, self.width))
            lo, hi = shard_range(eng.b_global, self.rank, self.world)
            eng.noise[s].copy_(n[lo:hi], non_blocking=True)

    # ------------------------------------------------------------------------------------------
    def _collect(self, eng: AdaptEngine, check_nan: bool = True) -> Tuple[Dict[Any, Tensor], Dict[str, Tensor]]:
        b = eng.b
        outputs: Dict[Any, Tensor] = {}
        arena = eng.out_arena.clone() if self.copy_outputs else None          # fresh tensors per call, like the reference
        for s in self.scales:
            outputs[('disp', s)] = eng.out_view(('disp', s), arena)
        for p, f in ((0, -1), (1, 1)):
            outputs[('axis_angle', 0, f)] = eng.axis_angle[p].reshape(b, 1, 3).clone()
            outputs[('translation', 0, f)] = eng.translation[p].reshape(b, 1, 3).clone()
            outputs[('cam_T_cam', 0, f)] = eng.cam_t_cam[p].clone()
        if self.emit_outputs:
            for s in self.scales:
                outputs[('depth', s)] = eng.out_view(('depth', s), arena)
                warped = eng.out_view(('warped', s), arena)
                for p, f in ((0, -1), (1, 1)):
                    outputs[('rgb', f, s)] = warped[p]
        vec = eng.losses.clone()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(vec)
        host = None
        if check_nan:
            host = vec.cpu()                            # the one sync per call (:1115)
            if self.losses_on_host:
                vec = host
        losses: Dict[str, Tensor] = {}
        slots = vec.unbind(0)                            # one call instead of 17 indexing ops: this runs after the sync, with the GPU idle
        for s in self.scales:
            for name in ('reprojection_loss', 'smooth_loss', 'reg_loss', 'depth_loss'):
                losses[f'{name}/scale_{s}'] = slots[LOSS_SLOTS[name] + s]
        total = vec[18:19]
        losses['depth_loss'] = total                     # same tensor as 'loss' 
        if eng.vel_scale > 0:
            losses['velocity_loss'] = slots[17]
        losses['loss'] = total
        if check_nan:
            total_host = float(host[18])
            if total_host != total_host:          # NaN (:1115)
                for k, v in losses.items():
                    print(k, float(v.reshape(-1)[0]))
                raise RuntimeError('NaN loss')
        return outputs, losses

    def _process_batch(self, inputs: Dict[Any, Tensor], loss_sample_weights: Optional[Tensor] = None,
                       use_online: bool = False) -> Tuple[Dict[Any, Tensor], Dict[str, Tensor]]:
        """Forward + losses of one mini-batch (:906-923).  Values of ``inputs`` are replaced by their
        device copies, as in the reference."""
        eng = self._stage_inputs(inputs, ns=self._set_of(use_online))          # online_models when asked (:930-932, :956-958)
        self._set_weights(eng, loss_sample_weights)
        self._draw_noise(eng)
        eng.forward()
        self._last_engine = eng
        return self._collect(eng)

    def backward(self, inputs_engine: Optional[AdaptEngine] = None) -> None:
        """Gradients of the last ``_process_batch`` into the flat gradient buffer (= ``loss.backward()``
        of :312 with the encoders frozen) followed by the data-parallel all-reduce."""
        eng = inputs_engine or self._last_engine
        eng.backward()
        self._allreduce_grads(eng)
        eng.net_set.optimizer.step_fn = lambda e=eng: self._apply_adam(e)          # `optimizer.step()` of the reference's loops

    @staticmethod
    def _allreduce_tensor(t: Tensor) -> None:
        import torch.distributed as dist
        dist.all_reduce(t)                               # SUM over the replicas (NCCL over NVLink / NVSwitch)

    def _allreduce_grads(self, eng: Optional[AdaptEngine] = None) -> None:
        """the exchange step of the data-parallel path.  With bucketed synchronisation (default) the engine's backward plan
        has already issued it, bucket by bucket, overlapped with the remaining weight-gradient kernels."""
        if self.world > 1 and not (eng is not None and eng.grad_sync_in_plan):
            self._allreduce_tensor(self._flat.grads)
            if eng is not None and getattr(eng, 'kind', 'adapt') != 'adapt':          # full training: the encoders' gradients too
                for ef in eng.net_set.enc_flat.values():
                    self._allreduce_tensor(ef.grads)

    def _apply_adam(self, eng) -> None:
        eng.adam(self._adam_lr())
        if getattr(eng, 'kind', 'adapt') != 'adapt':
            self.refresh_encoders()          # the kernel-layout copies of the (now changed) encoders are rebuilt on next use

    def gradients(self) -> Dict[Tuple[str, str], Tensor]:
        return self._flat.named_gradients()

    def encoder_gradients(self) -> Dict[Tuple[str, str], Tensor]:
        """gradients of the encoder tensors after a full-training backward pass: {(net, 'resnet.<torchvision key>'): tensor}"""
        out = {}
        for net, ef in self._main.enc_flat.items():
            if ef.grads is None:
                continue
            for key, g in ef.named_gradients().items():
                key = key.replace('.dsbn.', '.downsample.1.').replace('.ds.', '.downsample.0.')
                out[(net, 'resnet.' + key)] = g
        return out

    def _graph_ctx(self, g):
        """capture context of a step graph.  CLSLAM_LANE0_PRIORITY (default 0 = off): capture on a stream of that priority, so
        that the kernel nodes of lane 0 -- the critical chain of the step -- are scheduled ahead of the side lanes' when both
        have CTAs waiting for an SM (kernel nodes inherit the priority of the stream they were captured on)"""
        if LANE0_PRIORITY == 0:
            return torch.cuda.graph(g)
        if getattr(self, '_capture_stream', None) is None:
            self._capture_stream = torch.cuda.Stream(device=self.device, priority=LANE0_PRIORITY)
        return torch.cuda.graph(g, stream=self._capture_stream)

    def _adam_lr(self) -> float:
        return float(self.optimizer.param_groups[0]['lr'])

    def _step(self, eng: AdaptEngine) -> None:
        """one adaptation iteration on staged inputs: fwd, bwd, (all-reduce), Adam"""
        lr = self._adam_lr()
        eng.adam_hparams(lr)                 # read by the per-bucket Adam launches of a data-parallel plan (baked into its graph)
        ns = eng.net_set
        graphs = ns.graphs
        kind = getattr(eng, 'kind', 'adapt')
        n_run = ns.steps_run.get((eng.b, kind), 0)
        ns.steps_run[(eng.b, kind)] = n_run + 1
        late_event = getattr(self, '_late_event', None)
        if not self.use_cuda_graph or n_run == 0:       # first step eager: loads every kernel before capture
            if late_event is not None:
                torch.cuda.current_stream(self.device).wait_event(late_event)
            eng.forward_backward(); self._allreduce_grads(eng); eng.adam(lr)
            if kind != 'adapt':
                self.refresh_encoders()
            return
        if kind != 'adapt':
            # full training step: one graph on a single GPU; data-parallel: graph(fwd + bwd), all-reduce of the three flat
            # gradient buffers, graph(Adam)
            if late_event is not None:
                torch.cuda.current_stream(self.device).wait_event(late_event)
            if self.world == 1:
                key = (eng.b, kind, 'step', lr)
                if key not in graphs:
                    g = torch.cuda.CUDAGraph()
                    with self._graph_ctx(g):
                        eng.forward_backward(); eng.adam(lr)
                    graphs[key] = g
                graphs[key].replay()
            else:
                k1, k2 = (eng.b, kind, 'fwdbwd'), (eng.b, kind, 'adam', lr)
                if k1 not in graphs:
                    g1 = torch.cuda.CUDAGraph()
                    with self._graph_ctx(g1):
                        eng.forward_backward()
                    g2 = torch.cuda.CUDAGraph()
                    with self._graph_ctx(g2):
                        eng.adam(lr)
                    graphs[k1], graphs[k2] = g1, g2
                graphs[k1].replay(); self._allreduce_grads(eng); graphs[k2].replay()
            self.refresh_encoders()
            return
        if (self.world == 1 or eng.grad_sync_in_plan) and H2D_OVERLAP and late_event is not None:
            # two graphs: the networks (need rgb_aug / intrinsics only) run while the photometric inputs are still in flight
            # (with nothing in flight -- inputs already resident, second and later iterations of adapt(steps=k) -- the step
            # is the single graph below)
            ka, kb = (eng.b, 'nets'), (eng.b, 'rest', lr)
            if ka not in graphs:
                ga = torch.cuda.CUDAGraph()
                with self._graph_ctx(ga):
                    eng.networks_forward()
                graphs[ka] = ga
            if kb not in graphs:
                gb = torch.cuda.CUDAGraph()
                with self._graph_ctx(gb):
                    eng.photometric_and_backward(); eng.adam(lr)
                graphs[kb] = gb
            graphs[ka].replay()
            if late_event is not None:
                torch.cuda.current_stream(self.device).wait_event(late_event)
            graphs[kb].replay()
            return
        if self.world == 1 or eng.grad_sync_in_plan:
            # one graph: forward, backward (data-parallel: with the bucketed NCCL all-reduces inside), Adam
            key = (eng.b, 'step', lr)
            g = graphs.get(key)
            if g is None:
                g = torch.cuda.CUDAGraph()
                with self._graph_ctx(g):
                    eng.forward_backward(); eng.adam(lr)
                graphs[key] = g
            g.replay()
        else:
            k1, k2 = (eng.b, 'fwdbwd'), (eng.b, 'adam', lr)
            if k1 not in graphs:
                g1 = torch.cuda.CUDAGraph()
                with self._graph_ctx(g1):
                    eng.forward_backward()
                g2 = torch.cuda.CUDAGraph()
                with self._graph_ctx(g2):
                    eng.adam(lr)
                graphs[k1], graphs[k2] = g1, g2
            graphs[k1].replay()
            self._allreduce_grads(eng)
            graphs[k2].replay()

    # ------------------------------------------------------------------------------------------
    def adapt(self, online_data: Dict[Any, Tensor], training_data: Optional[Dict[Any, Tensor]] = None,
              online_index: int = 0, steps: int = 1, online_loss_weight: Optional[float] = None):
        """:291-319.  Returns (outputs, losses) of the LAST iteration (computed before its update)."""
        if online_loss_weight is None:
            loss_weights = None
        else:
            loss_weights = global_sample_weights(self.batch_size, online_loss_weight, online_index)
        if training_data is not None:
            self._set_adapt(freeze_encoder=True)
            eng = self._stage_inputs(training_data, defer_unused=True)
            self._set_weights(eng, loss_weights)
            for _ in range(steps):
                self._draw_noise(eng)
                self._step(eng)
                self._late_event = None     # the first iteration waited for the late inputs; the others find them resident
            self._flush_deferred()          # unused dict values move while the step runs
            self._last_engine = eng
            return self._collect(eng)
        self._set_eval()
        with torch.no_grad():
            return self._process_batch(online_data, loss_weights)

    def train(self, validate: bool = False, depth_error: bool = False, pose_error: bool = False, dataloader=None,
              verbose: bool = True, use_wandb: Optional[bool] = False) -> None:
        """:219-289.  The training loop over ``dataloader`` (batches in the dataset schema of SURVEY App. C.1): per batch
        zero_grad / _process_batch / backward / optimizer.step -- here one captured step of the full training plan --
        ``lr_scheduler.step()`` per epoch, checkpoints every ``save_frequency`` epochs.  Building the reference's own KITTI /
        Cityscapes loaders (``_create_dataloaders``, :829-904), validation, the depth / pose error tables and wandb logging
        are outside SURVEY.md section 8: pass the loader."""
        if use_wandb:
            raise NotImplementedError('wandb logging (:226-229, :1246-1267) is outside SURVEY.md section 8')
        if depth_error or pose_error:
            raise NotImplementedError('depth_error / pose_error need the ground truth of the dataset loaders: outside SURVEY.md section 8')
        if validate and self.val_loader is None:
            raise NotImplementedError('train(validate=True) needs `val_loader` to be assigned (see validate())')
        if dataloader is None:
            raise NotImplementedError('train() needs `dataloader`: the reference builds it from its datasets package '
                                      '(_create_dataloaders, :829-904), which is outside SURVEY.md section 8')
        self.train_loader = dataloader
        print(f'Training samples:   {len(self.train_loader):>5}')
        starting_epoch = self.epoch + 1
        for self.epoch in range(starting_epoch, self.num_epochs + 1):
            self._set_train()
            loss = []
            for sample_i in self.train_loader:
                eng = self._stage_inputs(sample_i)
                self._set_weights(eng, None)
                self._draw_noise(eng)
                self._step(eng)                              # zero_grad, forward, backward, optimizer.step (:251-256)
                self._last_engine = eng
                _, losses = self._collect(eng)
                loss.append(float(losses['loss']))
            if verbose and loss:
                print(f'Training epoch {self.epoch}/{self.num_epochs}: loss={sum(loss) / len(loss):.6f}')
            self.optimizer.torch_optimizer._opt_called = True          # the fused Adam has stepped (silences StepLR's order check)
            self.lr_scheduler.step()
            self.is_trained = True
            if self.save_frequency > 0 and self.epoch % self.save_frequency == 0:
                self.save_model()
            if validate:
                validation_loss = self.validate()          # :269-272
                if verbose:
                    print(f'Validation epoch {self.epoch}: loss={validation_loss:.6f}')
        if self.save_frequency > -1:
            self.save_model()

    # ------------------------------------------------------------------------------------------
    # inference entry points of the reference that Slam does not call (:530-626)
    def predict(self, batch) -> Dict[Any, Tensor]:
        """:530-536: eval-mode forward of one batch dict, returns the outputs"""
        if not self.is_trained:
            warnings.warn('The model has not been trained yet.', RuntimeWarning)
        self._set_eval()
        with torch.no_grad():
            outputs, _ = self._process_batch(batch)
        return outputs

    def _depth_of(self, image: Tensor) -> Tensor:
        """depth of `image` (n, 3, H, W) from the depth network alone (:547-550): encoder, decoder, dispconv_0, disp_to_depth on
        a plan of its own (no pose network, no losses, no collective; the adaptation engine's state is not touched).  Any
        H, W that are multiples of 32."""
        self._require_cuda()
        if image.dim() == 3:
            image = image.unsqueeze(0)
        n, _, h, w = image.shape
        ns = self._main
        prep = self._prepared('depth_encoder', ns)
        runner = ns.depth_runners.get((n, h, w))
        if runner is None:
            L.check(L.load().clslam_set_device(self.device.index or 0), 'set_device')
            runner = DepthRunner(prep, ns.flat, n, h, w, self.min_depth, self.max_depth, self.device,
                                 use_graph=self.use_cuda_graph)
            ns.depth_runners[(n, h, w)] = runner
        self._set_eval()
        return runner(image.to(self.device, dtype=torch.float32))

    def predict_from_image(self, image, as_numpy: bool = True):
        """:538-554: one image in, predicted depth out"""
        if not self.is_trained:
            warnings.warn('The model has not been trained yet.', RuntimeWarning)
        depth = self._depth_of(image)
        if as_numpy:
            depth = depth.squeeze().cpu().detach().numpy()
        return depth

    def predict_from_images(self, image_0: Tensor, image_1: Tensor, as_numpy: bool = True, return_loss: bool = False,
                            camera_matrix: Optional[Tensor] = None, inv_camera_matrix: Optional[Tensor] = None,
                            relative_distance: Optional[Tensor] = None):
        """:556-626: two images in, depth of both and the relative pose (image_0 -> image_1, not inverted) out"""
        if return_loss:
            raise NotImplementedError('predict_from_images(return_loss=True): the two-frame, single-scale loss of :603-624 is '
                                      'outside the adaptation hot path (SURVEY.md 8)')
        if not self.is_trained:
            warnings.warn('The model has not been trained yet.', RuntimeWarning)
        depth_0, depth_1 = self._depth_of(image_0), self._depth_of(image_1)
        was_trained, self.is_trained = self.is_trained, True          # the warning above was issued once
        try:
            transformation, _ = self.predict_pose(image_0, image_1, as_numpy=False)
        finally:
            self.is_trained = was_trained
        if as_numpy:
            depth_0 = depth_0.squeeze().cpu().detach().numpy()
            depth_1 = depth_1.squeeze().cpu().detach().numpy()
            transformation = transformation.squeeze().cpu().detach().numpy()
        return depth_0, depth_1, transformation

    def validate(self, dataloader=None) -> float:
        """:321-342: mean evaluation-mode loss over the validation loader (``dataloader=`` or a previously assigned
        ``val_loader``; building the reference's own loaders, :829-904, is outside SURVEY.md section 8).  ``save_val_depth`` plots
        are not produced."""
        if not self.is_trained:
            warnings.warn('The model has not been trained yet.', RuntimeWarning)
        if dataloader is not None:
            self.val_loader = dataloader
        if self.val_loader is None:
            raise NotImplementedError('validate() needs `dataloader` (or `val_loader`): the reference builds it from its datasets '
                                      'package (_create_dataloaders, :829-904), which is outside SURVEY.md section 8')
        self._set_eval()
        loss = []
        with torch.no_grad():
            for sample_i in self.val_loader:
                _, losses = self._process_batch(sample_i)
                loss.append(float(losses['loss']))
        return float(np.mean(loss)) if loss else float('nan')

    def compute_depth_error(self, median_scaling: bool = True, print_results: bool = True) -> Dict[str, float]:
        raise NotImplementedError('compute_depth_error() (:344-468) needs the dataset loaders: outside SURVEY.md 8')

    def compute_pose_error(self, print_results: bool = True) -> Dict[str, float]:
        raise NotImplementedError('compute_pose_error() (:470-528) needs the dataset loaders: outside SURVEY.md 8')

    def save_prediction(self, inputs: Dict[Any, Tensor], outputs: Dict[Any, Tensor]) -> None:
        raise NotImplementedError('save_prediction() (:1197 ff.) is plotting (matplotlib): outside SURVEY.md 8')

    def predict_pose(self, image_0: Tensor, image_1: Tensor, as_numpy: bool = True, use_online: bool = False):
        """:628-664: pose net forward for a loop-closure pair, invert=False; returns (4x4, eye(6))."""
        if not self.is_trained:
            warnings.warn('The model has not been trained yet.', RuntimeWarning)
        ns = self._set_of(use_online)            # :650-652
        self._require_cuda()
        if image_0.dim() == 3:
            image_0 = image_0.unsqueeze(0)
        if image_1.dim() == 3:
            image_1 = image_1.unsqueeze(0)
        self._set_eval()
        runner = self._get_pose_runner(image_0.shape[0], image_0.shape[2], image_0.shape[3], ns)
        transformation = runner(image_0.to(self.device, dtype=torch.float32), image_1.to(self.device, dtype=torch.float32))
        cov_matrix = torch.eye(6, device=self.device)
        if as_numpy:
            return transformation.squeeze().cpu().numpy(), cov_matrix.cpu().numpy()
        return transformation, cov_matrix

    def _get_pose_runner(self, n: int, h: int, w: int, ns: Optional[_NetSet] = None):
        ns = ns or self._main
        prep = self._prepared('pose_encoder', ns)
        r = ns.pose_runner
        if r is None or (r.n, r.h, r.w) != (n, h, w):
            L.check(L.load().clslam_set_device(self.device.index or 0), 'set_device')
            r = _PoseRunner(prep, ns.flat, n, h, w, self.device, use_graph=self.use_cuda_graph)
            ns.pose_runner = r
        return r

    # ------------------------------------------------------------------------------------------
    # checkpoints (:669-749): same files, key names and shapes as the reference
    def save_model(self) -> None:
        save_folder = Path(self.log_path) / 'models' / f'weights_{self.epoch:03}'
        if self.rank != 0:          # data-parallel replicas hold bit-identical weights: one writer
            return
        save_folder.mkdir(parents=True, exist_ok=True)
        for model_name, model in self.models.items():
            to_save = {k: v.detach().clone().contiguous() for k, v in model.state_dict().items()}
            if 'encoder' in model_name:
                # :684-688 writes `Tensor(self.height)` / `Tensor(self.width)`: UNINITIALISED tensors of that many elements
                # ("the sizes needed at prediction time" of monodepth2-style loaders are their lengths); zeros here
                to_save['height'] = torch.zeros(self.height)
                to_save['width'] = torch.zeros(self.width)
            torch.save(to_save, save_folder / f'{model_name}.pth')
        torch.save({'optimizer': self.optimizer.state_dict(), 'scheduler': self.lr_scheduler.state_dict()},
                   save_folder / 'optimizer.pth')
        if self.config_file is not None and Path(self.config_file).is_file():
            shutil.copy(self.config_file, Path(self.log_path) / 'config.yaml')
        print(f'Saved model to: {save_folder}')

    def load_model(self, load_optimizer: bool = True) -> None:
        if self.load_weights_folder is None:
            print('Weights folder required to load the model is not specified.')
        if not self.load_weights_folder.exists():
            print(f'Cannot find folder: {self.load_weights_folder}')
        print(f'Load model from: {self.load_weights_folder}')
        for model_name, model in self.models.items():
            path = self.load_weights_folder / f'{model_name}.pth'
            pretrained = torch.load(path, map_location=self.device)
            own = model.state_dict()
            pretrained = {k: v for k, v in pretrained.items() if k in own}      # drops 'height'/'width'
            if len(pretrained) == 0:
                raise RuntimeError(f'No fitting weights found in: {path}')
            own.update(pretrained)
            model.load_state_dict(own)
        self.refresh_encoders()
        self.is_trained = True
        if load_optimizer:
            try:
                optimizer_dict = torch.load(self.load_weights_folder / 'optimizer.pth', map_location=self.device)
                if 'optimizer' in optimizer_dict:
                    self.optimizer.load_state_dict(optimizer_dict['optimizer'])
                    self.lr_scheduler.load_state_dict(optimizer_dict['scheduler'])
                    self.epoch = self.lr_scheduler.last_epoch
                    print(f'Restored optimizer and LR scheduler (resume from epoch {self.epoch}).')
                else:
                    self.optimizer.load_state_dict(optimizer_dict)
                    print('Restored optimizer (legacy mode).')
            except Exception:   # noqa: BLE001  (the reference swallows everything here, :747)
                print('Cannot find matching optimizer weights, so the optimizer is randomly initialized.')

    def load_online_model(self, load_optimizer: bool = True) -> None:
        """:751-792: the checkpoint into the online copy of the networks (and its optimiser)"""
        if self.load_weights_folder is None:
            print('Weights folder required to load the model is not specified.')
        if not self.load_weights_folder.exists():
            print(f'Cannot find folder: {self.load_weights_folder}')
        print(f'Load online model from: {self.load_weights_folder}')
        for model_name, model in self.online_models.items():
            path = self.load_weights_folder / f'{model_name}.pth'
            pretrained = torch.load(path, map_location=self.device)
            own = model.state_dict()
            pretrained = {k: v for k, v in pretrained.items() if k in own}
            if len(pretrained) == 0:
                raise RuntimeError(f'No fitting weights found in: {path}')
            own.update(pretrained)
            model.load_state_dict(own)
        self.refresh_encoders()
        if load_optimizer:
            try:
                optimizer_dict = torch.load(self.load_weights_folder / 'optimizer.pth', map_location=self.device)
                if 'optimizer' in optimizer_dict:
                    self.online_optimizer.load_state_dict(optimizer_dict['optimizer'])
                    print('Restored online optimizer')
                else:
                    self.online_optimizer.load_state_dict(optimizer_dict)
                    print('Restored online optimizer (legacy mode).')
            except Exception:   # noqa: BLE001  (:789)
                print('Cannot find matching optimizer weights, so the optimizer is randomly initialized.')


class _PoseRunner:
    """pose encoder + decoder + head for ``predict_pose`` (batch n, both 'passes' carry the same pair)."""

    def __init__(self, prep: PreparedEncoder, flat: FlatParams, n: int, h: int, w: int, device, use_graph: bool = True):
        from .engine import GraphedPlan, build_encoder
        self.n, self.h, self.w = n, h, w
        plan = Plan(device)
        lib = plan.lib
        self.img0 = plan.alloc(n, 3, h, w)
        self.img1 = plan.alloc(n, 3, h, w)
        from .engine import stage_for_encoder
        staged, s2d = stage_for_encoder(plan, prep, 'stage', [(self.img0, self.img1, 0, n), (self.img0, self.img1, n, n)], 2 * n, h, w)
        f4 = build_encoder(plan, prep, staged, 2 * n, h, w, 'pose_enc', x_s2d=s2d)[4]
        hp, wp = f4.shape[1], f4.shape[2]
        x = f4
        for name, co, ci, k, act in (('squeeze', 256, 512, 1, L.ACT_RELU), ('pose_0', 256, 256, 3, L.ACT_RELU),
                                     ('pose_1', 256, 256, 3, L.ACT_RELU), ('pose_2', 12, 256, 1, L.ACT_NONE)):
            d = plan.desc(2 * n, hp, wp, ci, co, k, 1, (k - 1) // 2, act=act)
            y = plan.alloc(2 * n, hp, wp, co)
            plan.conv_fwd(name, d, x, None, flat.weight_ptr(('pose_decoder', name)), None,
                          flat.bias_ptr(('pose_decoder', name)), None, y)      # fp32 weights (no bf16 planes needed here)
            x = y
        self.k_eye = torch.eye(4, device=device).repeat(n, 1, 1).contiguous()
        self.aa = plan.alloc(2, n, 3); self.tr = plan.alloc(2, n, 3)
        self.cam = plan.alloc(2, n, 4, 4); self.proj = plan.alloc(2, n, 12)
        plan.simple('pose_head', lib.clslam_pose_head_fwd, _p(x), _p(self.k_eye), n, hp, wp, 0, _p(self.aa), _p(self.tr),
                    _p(self.cam), _p(self.proj))
        plan.finalize()
        self.plan = plan
        self.graphed = GraphedPlan(plan, use_graph)

    def __call__(self, image_0: Tensor, image_1: Tensor) -> Tensor:
        self.img0.copy_(image_0); self.img1.copy_(image_1)
        self.graphed.run()
        return self.cam[0].clone()
