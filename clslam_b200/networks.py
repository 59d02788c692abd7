"""Parameter containers with the reference's module/attribute names, so that state dicts
(checkpoints) are key- and shape-compatible with the reference:

    ResnetEncoder  -> 'resnet.conv1.weight', 'resnet.layer2.0.downsample.1.running_var', 'resnet.fc.bias', ...
                      (networks/resnet_encoder.py:79-113; torchvision ResNet-18 naming)
    DepthDecoder   -> 'upconv_4_0.conv.conv.weight', 'dispconv_0.conv.bias', ...
                      (networks/depth_decoder.py:30-46, networks/layers.py:9-48)
    PoseDecoder    -> 'squeeze.weight', 'pose_0.weight', ...  (networks/pose_decoder.py:27-30)

These classes hold tensors only; all arithmetic runs in the CUDA kernels driven by
``clslam_b200.engine``.  ``forward`` of the encoder is wired by the predictor because
``slam/slam.py:146`` calls ``predictor.models['depth_encoder'](image)`` directly.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

import torch
from torch import nn

DEC_CH = (16, 32, 64, 128, 256)       # networks/depth_decoder.py:28
ENC_CH = (64, 64, 128, 256, 512)      # networks/resnet_encoder.py:99


class _Conv(nn.Module):
    """weight (+bias) holder; init follows nn.Conv2d defaults unless overwritten by a checkpoint."""

    def __init__(self, cin: int, cout: int, k: int, bias: bool, device=None):
        super().__init__()
        w = torch.empty(cout, cin, k, k, device=device)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w)
        if bias:
            bound = 1 / math.sqrt(cin * k * k)
            self.bias = nn.Parameter(torch.empty(cout, device=device).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)


class _BatchNorm(nn.Module):
    def __init__(self, c: int, device=None):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c, device=device))
        self.bias = nn.Parameter(torch.zeros(c, device=device))
        self.register_buffer('running_mean', torch.zeros(c, device=device))
        self.register_buffer('running_var', torch.ones(c, device=device))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long, device=device))
        self.eps = 1e-5


class _BasicBlock(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int, device=None):
        super().__init__()
        self.conv1 = _Conv(cin, cout, 3, False, device)
        self.bn1 = _BatchNorm(cout, device)
        self.conv2 = _Conv(cout, cout, 3, False, device)
        self.bn2 = _BatchNorm(cout, device)
        self.stride = stride
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(_Conv(cin, cout, 1, False, device), _BatchNorm(cout, device))
        else:
            self.downsample = None


class _Linear(nn.Module):
    """resnet.fc: present in the reference's state dict (and in its Adam param list), never used."""

    def __init__(self, cin: int, cout: int, device=None):
        super().__init__()
        bound = 1 / math.sqrt(cin)
        self.weight = nn.Parameter(torch.empty(cout, cin, device=device).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(cout, device=device).uniform_(-bound, bound))


class _ResNet18Trunk(nn.Module):
    def __init__(self, in_ch: int, device=None):
        super().__init__()
        self.conv1 = _Conv(in_ch, 64, 7, False, device)
        self.bn1 = _BatchNorm(64, device)
        cin = 64
        for li, planes in enumerate((64, 128, 256, 512), start=1):
            blocks = []
            for bi in range(2):
                blocks.append(_BasicBlock(cin, planes, 2 if (li > 1 and bi == 0) else 1, device))
                cin = planes
            setattr(self, f'layer{li}', nn.Sequential(*blocks))
        self.fc = _Linear(512, 1000, device)
        for m in self.modules():                       # torchvision / resnet_encoder.py:40-45 init
            if isinstance(m, _Conv):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')


class ResnetEncoder(nn.Module):
    """ResNet-18 encoder (num_layers=18 only; resnet_encoder.py:79).  ``num_input_images=2`` is the
    6-channel pose encoder (:25-30)."""

    def __init__(self, num_layers: int, pretrained: bool, num_input_images: int = 1, device=None):
        super().__init__()
        if num_layers != 18:
            raise ValueError(f'clslam_b200 implements ResNet-18 encoders only (got {num_layers}).')
        if num_input_images < 1:
            raise ValueError(f'Invalid value ({num_input_images}) for num_input_images.')
        if pretrained:
            raise ValueError('resnet_pretrained=True needs a network download; load weights with load_model().')
        self.num_ch_encoder = list(ENC_CH)
        self.num_input_images = num_input_images
        self.resnet = _ResNet18Trunk(3 * num_input_images, device)
        self._runner: Optional[Callable[[torch.Tensor], List[torch.Tensor]]] = None
        self.version = 0          # bumped whenever weights may have changed (invalidates prepared copies)

    def bind(self, runner: Callable[[torch.Tensor], List[torch.Tensor]]) -> None:
        self._runner = runner

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        if self._runner is None:
            raise RuntimeError('encoder is not bound to a CUDA engine')
        return self._runner(x)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.version += 1
        return out


class _ConvBlockHolder(nn.Module):
    """networks/layers.py:9-25 ConvBlock = Conv3x3 (+ELU): attribute path '.conv.conv.weight'."""

    def __init__(self):
        super().__init__()
        self.conv = _Conv3x3Holder()


class _Conv3x3Holder(nn.Module):
    """networks/layers.py:28-48 Conv3x3: attribute path '.conv.weight'."""

    def __init__(self):
        super().__init__()
        self.conv = _FlatConv()


class _FlatConv(nn.Module):
    """weight/bias are views into the predictor's flat parameter buffer (set by FlatParams)."""

    def __init__(self):
        super().__init__()
        self.register_parameter('weight', None)
        self.register_parameter('bias', None)

    def _apply(self, fn, recurse=True):
        """``model.to(...)`` / ``.cuda()`` / ``.float()`` as a no-op is fine (the reference calls ``m.to(self.device)``, :186);
        a conversion that would move or re-type the tensors is refused: the kernels, Adam and the all-reduce work on the flat
        buffer these Parameters are views of, and a converted copy would silently detach them from it."""
        for name in ('weight', 'bias'):
            p = getattr(self, name)
            if p is None:
                continue
            q = fn(p.data)
            if q.device != p.device or q.dtype != p.dtype:
                raise RuntimeError(f'clslam_b200: the decoder parameters are views of the flat parameter buffer on {p.device} '
                                   f'({p.dtype}); moving / converting them ({q.device}, {q.dtype}) is not supported')
        return self


def depth_decoder_spec() -> List[Tuple[str, int, int, int]]:
    """(module path, cout, cin, k) in the reference's creation order (depth_decoder.py:30-46)."""
    spec = []
    for i in range(4, -1, -1):
        cin0 = ENC_CH[-1] if i == 4 else DEC_CH[i + 1]
        spec.append((f'upconv_{i}_0.conv.conv', DEC_CH[i], cin0, 3))
        cin1 = DEC_CH[i] + (ENC_CH[i - 1] if i > 0 else 0)
        spec.append((f'upconv_{i}_1.conv.conv', DEC_CH[i], cin1, 3))
    for s in range(4):
        spec.append((f'dispconv_{s}.conv', 1, DEC_CH[s], 3))
    return spec


def pose_decoder_spec() -> List[Tuple[str, int, int, int]]:
    return [('squeeze', 256, 512, 1), ('pose_0', 256, 256, 3), ('pose_1', 256, 256, 3), ('pose_2', 12, 256, 1)]


class DepthDecoder(nn.Module):
    def __init__(self, num_ch_encoder=ENC_CH, scales=(0, 1, 2, 3)):
        super().__init__()
        if tuple(scales) != (0, 1, 2, 3):
            raise ValueError('clslam_b200 implements scales=(0, 1, 2, 3) only')
        self.scales = tuple(scales)
        for i in range(4, -1, -1):
            setattr(self, f'upconv_{i}_0', _ConvBlockHolder())
            setattr(self, f'upconv_{i}_1', _ConvBlockHolder())
        for s in self.scales:
            setattr(self, f'dispconv_{s}', _Conv3x3Holder())


class PoseDecoder(nn.Module):
    def __init__(self):
        super().__init__()
        for name, _, _, _ in pose_decoder_spec():
            setattr(self, name, _FlatConv())


class FlatParams:
    """One flat fp32 buffer holding the 36 decoder tensors that ``adapt`` trains, with weights stored
    [Cout][kh][kw][Cin] (the kernels' layout).  The nn.Parameters handed to the modules are
    *views* with the reference's logical (Cout,Cin,kh,kw) shape, so ``state_dict()`` /
    ``load_state_dict()`` speak the reference's checkpoint format while Adam and the wgrad kernels
    work on the flat buffer in place."""

    def __init__(self, depth_decoder: DepthDecoder, pose_decoder: PoseDecoder, device):
        entries = []      # (net, path, cout, cin, k)
        for path, co, ci, k in depth_decoder_spec():
            entries.append(('depth_decoder', path, co, ci, k))
        for path, co, ci, k in pose_decoder_spec():
            entries.append(('pose_decoder', path, co, ci, k))
        # every tensor starts on a 256-byte boundary (float4 / TMA alignment); the padding elements stay 0
        def aligned(n):
            return (n + 63) // 64 * 64
        total = sum(aligned(co * ci * k * k) + aligned(co) for _, _, co, ci, k in entries)
        self.numel = total
        self.numel_logical = sum(co * ci * k * k + co for _, _, co, ci, k in entries)
        self.params = torch.zeros(total, device=device)
        self.grads = torch.zeros(total, device=device)
        self.exp_avg = torch.zeros(total, device=device)
        self.exp_avg_sq = torch.zeros(total, device=device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=device)   # Adam updates done so far
        # bf16 hi / lo planes of `params` (refreshed on the device at the start of every forward pass)
        # (one [2][total] allocation: the lo plane follows the hi plane, so a single TMA map covers both)
        self.params_hl = torch.zeros(2, total, dtype=torch.bfloat16, device=device)
        self.params_hi = self.params_hl[0]
        self.params_lo = self.params_hl[1]
        self.index = {}   # (net, path) -> dict(w_off, w_numel, b_off, co, ci, k)
        off = 0
        mods = {'depth_decoder': depth_decoder, 'pose_decoder': pose_decoder}
        for net, path, co, ci, k in entries:
            wn = co * ci * k * k
            info = dict(w_off=off, w_numel=wn, b_off=off + aligned(wn), co=co, ci=ci, k=k)
            self.index[(net, path)] = info
            holder = mods[net].get_submodule(path)
            w_view = self.view_weight(self.params, info)
            b_view = self.params[info['b_off']:info['b_off'] + co]
            bound = 1 / math.sqrt(ci * k * k)
            with torch.no_grad():
                w_view.uniform_(-bound, bound)
                b_view.uniform_(-bound, bound)
            holder.weight = nn.Parameter(w_view)
            holder.bias = nn.Parameter(b_view)
            holder.weight.grad = self.view_weight(self.grads, info)
            holder.bias.grad = self.grads[info['b_off']:info['b_off'] + co]
            off += aligned(wn) + aligned(co)

    @staticmethod
    def view_weight(flat: torch.Tensor, info) -> torch.Tensor:
        """logical (Cout, Cin, kh, kw) view of the [Cout][kh][kw][Cin] storage"""
        co, ci, k = info['co'], info['ci'], info['k']
        return flat[info['w_off']:info['w_off'] + info['w_numel']].view(co, k, k, ci).permute(0, 3, 1, 2)

    def weight_ptr(self, key) -> int:
        return self.params.data_ptr() + 4 * self.index[key]['w_off']

    def weight_hi_ptr(self, key) -> int:
        return self.params_hi.data_ptr() + 2 * self.index[key]['w_off']

    def weight_lo_ptr(self, key) -> int:
        return self.params_lo.data_ptr() + 2 * self.index[key]['w_off']

    def bias_ptr(self, key) -> int:
        return self.params.data_ptr() + 4 * self.index[key]['b_off']

    def wgrad_ptr(self, key) -> int:
        return self.grads.data_ptr() + 4 * self.index[key]['w_off']

    def bgrad_ptr(self, key) -> int:
        return self.grads.data_ptr() + 4 * self.index[key]['b_off']

    def named_gradients(self):
        """{(net, 'path.weight'|'path.bias'): tensor} in the reference's logical shapes"""
        out = {}
        for (net, path), info in self.index.items():
            out[(net, path + '.weight')] = self.view_weight(self.grads, info)
            out[(net, path + '.bias')] = self.grads[info['b_off']:info['b_off'] + info['co']]
        return out


class EncoderFlat:
    """The trainable tensors of one ResNet-18 encoder in flat fp32 buffers (full training path):
    convolution weights stored [Cout][kh][kw][Cin] (the kernels' layout; conv1's input channels zero-padded to 4 / 8 like the
    staged images), BatchNorm weight / bias, plus gradient and Adam-moment buffers of the same layout.  The module's
    nn.Parameters are re-bound to *views* with the reference's logical shapes, so ``state_dict()`` / ``load_state_dict()`` /
    the optimiser's parameter list keep speaking the reference's format (resnet_encoder.py:79-113, torchvision key names) while
    the wgrad kernels, Adam and the all-reduce work on the flat buffers in place.  ``resnet.fc`` (never used, never trained)
    and the BatchNorm running statistics stay ordinary module tensors."""

    def __init__(self, enc: ResnetEncoder, device):
        self.enc = enc
        r = enc.resnet
        cin_pad = 4 if enc.num_input_images == 1 else 8
        self.cin_pad = cin_pad
        items = []          # (name, module, kind, shape info)
        items.append(('conv1', r.conv1, 'conv', (64, 3 * enc.num_input_images, 7, cin_pad)))
        items.append(('bn1', r.bn1, 'bn', 64))
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(r, f'layer{li}')):
                p = f'layer{li}.{bi}'
                for cname in ('conv1', 'conv2'):
                    conv = getattr(blk, cname)
                    co, ci, k, _ = conv.weight.shape
                    items.append((f'{p}.{cname}', conv, 'conv', (co, ci, k, ci)))
                    items.append((f'{p}.bn{cname[-1]}', getattr(blk, 'bn' + cname[-1]), 'bn', co))
                if blk.downsample is not None:
                    conv = blk.downsample[0]
                    co, ci, k, _ = conv.weight.shape
                    items.append((f'{p}.ds', conv, 'conv', (co, ci, k, ci)))
                    items.append((f'{p}.dsbn', blk.downsample[1], 'bn', co))

        def aligned(n):
            return (n + 63) // 64 * 64
        total = 0
        self.index = {}
        for name, mod, kind, info in items:
            if kind == 'conv':
                co, ci, k, cs = info
                self.index[name] = dict(kind='conv', off=total, co=co, ci=ci, k=k, cs=cs, numel=co * k * k * cs)
                total += aligned(co * k * k * cs)
            else:
                self.index[name] = dict(kind='bn', off=total, c=info, numel=2 * info)
                total += aligned(2 * info)
        self.numel = total
        self.params = torch.zeros(total, device=device)
        # gradient and Adam-moment buffers (3 x 45 MB per encoder) only exist once the encoder is trained: the per-frame
        # adaptation path never touches them
        self.grads: Optional[torch.Tensor] = None
        self.exp_avg: Optional[torch.Tensor] = None
        self.exp_avg_sq: Optional[torch.Tensor] = None
        self.modules = {}
        with torch.no_grad():
            for name, mod, kind, info in items:
                ix = self.index[name]
                self.modules[name] = mod
                if kind == 'conv':
                    old = mod.weight.detach().to(device)
                    view = self.view_weight(self.params, ix)
                    view.copy_(old)
                    mod.weight = nn.Parameter(view)
                else:
                    c = ix['c']
                    wv, bv = self.params[ix['off']:ix['off'] + c], self.params[ix['off'] + c:ix['off'] + 2 * c]
                    wv.copy_(mod.weight.detach().to(device)); bv.copy_(mod.bias.detach().to(device))
                    mod.weight = nn.Parameter(wv)
                    mod.bias = nn.Parameter(bv)

    def ensure_training_buffers(self) -> None:
        if self.grads is not None:
            return
        dev = self.params.device
        self.grads = torch.zeros(self.numel, device=dev)
        self.exp_avg = torch.zeros(self.numel, device=dev)
        self.exp_avg_sq = torch.zeros(self.numel, device=dev)
        for name, ix in self.index.items():
            mod = self.modules[name]
            if ix['kind'] == 'conv':
                mod.weight.grad = self.view_weight(self.grads, ix)
            else:
                c = ix['c']
                mod.weight.grad = self.grads[ix['off']:ix['off'] + c]
                mod.bias.grad = self.grads[ix['off'] + c:ix['off'] + 2 * c]

    def state_triples(self):
        """[(Parameter, exp_avg view, exp_avg_sq view)] in module order, for the optimiser's state dict"""
        out = []
        if self.exp_avg is None:
            return out
        for name, ix in self.index.items():
            mod = self.modules[name]
            if ix['kind'] == 'conv':
                out.append((mod.weight, self.view_weight(self.exp_avg, ix), self.view_weight(self.exp_avg_sq, ix)))
            else:
                c, o = ix['c'], ix['off']
                out.append((mod.weight, self.exp_avg[o:o + c], self.exp_avg_sq[o:o + c]))
                out.append((mod.bias, self.exp_avg[o + c:o + 2 * c], self.exp_avg_sq[o + c:o + 2 * c]))
        return out

    @staticmethod
    def view_weight(flat: torch.Tensor, ix) -> torch.Tensor:
        """logical (Cout, Cin, kh, kw) view of the [Cout][kh][kw][Cin_stored] storage"""
        co, ci, k, cs = ix['co'], ix['ci'], ix['k'], ix['cs']
        return flat[ix['off']:ix['off'] + ix['numel']].view(co, k, k, cs)[..., :ci].permute(0, 3, 1, 2)

    def w_ptr(self, name: str) -> int:
        return self.params.data_ptr() + 4 * self.index[name]['off']

    def dw_ptr(self, name: str) -> int:
        return self.grads.data_ptr() + 4 * self.index[name]['off']

    def bn_ptrs(self, name: str):
        """(gamma, beta, d gamma, d beta) device addresses"""
        ix = self.index[name]
        p, g = self.params.data_ptr() + 4 * ix['off'], self.grads.data_ptr() + 4 * ix['off']
        return p, p + 4 * ix['c'], g, g + 4 * ix['c']

    def frozen_mask(self) -> torch.Tensor:
        """1.0 where the element's Parameter has requires_grad, else 0.0 (the reference's optimiser skips parameters without a
        gradient: _set_adapt(freeze_encoder=False) leaves the 'bn'-named BatchNorm parameters frozen, :813-816)"""
        m = torch.zeros(self.numel, device=self.params.device)
        for name, ix in self.index.items():
            mod = self.modules[name]
            if ix['kind'] == 'conv':
                if mod.weight.requires_grad:
                    m[ix['off']:ix['off'] + ix['numel']] = 1.0
            else:
                c = ix['c']
                if mod.weight.requires_grad:
                    m[ix['off']:ix['off'] + c] = 1.0
                if mod.bias.requires_grad:
                    m[ix['off'] + c:ix['off'] + 2 * c] = 1.0
        return m

    def named_gradients(self):
        out = {}
        for name, ix in self.index.items():
            if ix['kind'] == 'conv':
                out[name + '.weight'] = self.view_weight(self.grads, ix)
            else:
                c = ix['c']
                out[name + '.weight'] = self.grads[ix['off']:ix['off'] + c]
                out[name + '.bias'] = self.grads[ix['off'] + c:ix['off'] + 2 * c]
        return out
