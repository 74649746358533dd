"""Synthetic workloads of BASELINE.json: the networks whose Hebbian layers the benchmark drives.

The reference's model files cannot travel to the GPU box, so the two topologies the north-star
names are restated here from stock torch.nn modules with the SAME module tree / parameter names
(so exclude lists and checkpoints carry over):
  unet2d()  == models/networks_2d/unet.py: UNet_Transposed_Leaky (:423-478) built by unet() (:705-708)
  unet3d()  == models/networks_3d/unet3d.py: UNet3D (:31-126) built by unet3d() (:226-229)
tests/test_workloads.py checks the module trees against fixtures taken from the reference.
These are host-side shape sources only — every Conv/ConvTranspose in them is swapped for a
Hebbian layer by makehebbian() (oracle or CUDA drop-in) before use.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

GLAS_MEAN = [0.787803, 0.512017, 0.784938]     # config/dataset_config/dataset_cfg.py:11-12
GLAS_STD = [0.428206, 0.507778, 0.426366]
EXCLUDE_2D = ['out_conv_dp1', 'out_conv_dp2', 'out_conv_dp3', 'out_conv']          # reproduce_..._2d.sh:40
EXCLUDE_3D = ['conv', 'dsv1', 'dsv2', 'dsv3', 'dsv4', 'out_conv', 'out_sdf', 'out_seg']  # ..._3d.sh:41


def _pair_of_convs(cin, cout, act, p_drop=None):
    layers = [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), act()]
    if p_drop is not None:
        layers.append(nn.Dropout(p_drop))
    layers += [nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), act()]
    return nn.Sequential(*layers)


class _Named(nn.Module):
    """A module holding one child under a given attribute name and forwarding to it."""

    def __init__(self, name, child):
        super().__init__()
        self._fwd = name
        self.add_module(name, child)

    def forward(self, x):
        return getattr(self, self._fwd)(x)


class _Enc2d(nn.Module):
    def __init__(self, cin, widths, drops):
        super().__init__()
        self.in_conv = _Named('conv_conv', _pair_of_convs(cin, widths[0], nn.LeakyReLU, drops[0]))
        for i in range(1, 5):
            block = _Named('conv_conv', _pair_of_convs(widths[i - 1], widths[i], nn.LeakyReLU, drops[i]))
            self.add_module(f'down{i}', _Named('maxpool_conv', nn.Sequential(nn.MaxPool2d(2), block)))

    def forward(self, x):
        feats = [self.in_conv(x)]
        for i in range(1, 5):
            feats.append(getattr(self, f'down{i}')(feats[-1]))
        return feats


class _Up2d(nn.Module):
    def __init__(self, c_low, c_skip, c_out):
        super().__init__()
        self.conv1x1 = nn.Conv2d(c_low, c_skip, kernel_size=1)
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.conv = _Named('conv', _pair_of_convs(c_skip * 2, c_out, lambda: nn.ReLU(inplace=True)))

    def forward(self, low, skip):
        return self.conv(torch.cat([skip, self.up(self.conv1x1(low))], dim=1))


class _Dec2d(nn.Module):
    def __init__(self, widths):
        super().__init__()
        for i in range(1, 5):
            self.add_module(f'up{i}', _Up2d(widths[5 - i], widths[4 - i], widths[4 - i]))

    def forward(self, feats):
        x = feats[4]
        for i in range(1, 5):
            x = getattr(self, f'up{i}')(x, feats[4 - i])
        return x


class UNet2D(nn.Module):
    def __init__(self, in_chns=3, class_num=2, widths=(16, 32, 64, 128, 256), drops=(0.05, 0.1, 0.2, 0.3, 0.5)):
        super().__init__()
        self.encoder = _Enc2d(in_chns, widths, drops)
        self.main_decoder = _Dec2d(widths)
        w0 = widths[0]
        self.out_conv = nn.Sequential(
            nn.Conv2d(w0, w0 * 4, kernel_size=3, padding=1), nn.ReLU(), nn.Dropout(),
            nn.Conv2d(w0 * 4, w0 * 2, kernel_size=3, padding=1), nn.ReLU(), nn.Dropout(),
            nn.Conv2d(w0 * 2, class_num, kernel_size=3, padding=1))

    def forward(self, x):
        return self.out_conv(self.main_decoder(self.encoder(x)))


def init_weights_like_reference(net, gain=0.02):
    """init_weights(net,'kaiming') of models/networks_2d/unet.py:7-28 (also hits Hebbian layers:
    their class names contain 'Conv')."""
    for m in net.modules():
        name = m.__class__.__name__
        if hasattr(m, 'weight') and m.weight is not None and ('Conv' in name or 'Linear' in name):
            nn.init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            if getattr(m, 'bias', None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif 'BatchNorm2d' in name:
            nn.init.normal_(m.weight.data, 1.0, gain)
            nn.init.constant_(m.bias.data, 0.0)
    return net


def unet2d(in_channels=3, num_classes=2):
    return init_weights_like_reference(UNet2D(in_channels, num_classes))


def _block3d(cin, f, name):
    return nn.Sequential(OrderedDict([
        (name + 'conv1', nn.Conv3d(cin, f, kernel_size=3, padding=1, bias=True)),
        (name + 'norm1', nn.BatchNorm3d(f)), (name + 'relu1', nn.ReLU(inplace=True)),
        (name + 'conv2', nn.Conv3d(f, f, kernel_size=3, padding=1, bias=True)),
        (name + 'norm2', nn.BatchNorm3d(f)), (name + 'relu2', nn.ReLU(inplace=True))]))


class UNet3D(nn.Module):
    def __init__(self, in_channels=1, out_channels=2, init_features=64):
        super().__init__()
        f = init_features
        chans = [in_channels, f, 2 * f, 4 * f, 8 * f]
        for i in range(1, 5):
            self.add_module(f'encoder{i}', _block3d(chans[i - 1], chans[i], f'enc{i}'))
            self.add_module(f'pool{i}', nn.MaxPool3d(kernel_size=2, stride=2))
        self.bottleneck = _block3d(8 * f, 16 * f, 'bottleneck')
        for i in range(4, 0, -1):
            self.add_module(f'upconv{i}', nn.ConvTranspose3d(chans[i] * 2, chans[i], kernel_size=2, stride=2))
            self.add_module(f'decoder{i}', _block3d(chans[i] * 2, chans[i], f'dec{i}'))
        self.conv = nn.Conv3d(f, out_channels, kernel_size=1)

    def forward(self, x):
        skips = []
        for i in range(1, 5):
            x = getattr(self, f'encoder{i}')(x)
            skips.append(x)
            x = getattr(self, f'pool{i}')(x)
        x = self.bottleneck(x)
        for i in range(4, 0, -1):
            x = getattr(self, f'upconv{i}')(x)
            x = getattr(self, f'decoder{i}')(torch.cat((x, skips[i - 1]), dim=1))
        return self.conv(x)


def unet3d(in_channels=1, num_classes=2, init_features=64):
    return init_weights_like_reference(UNet3D(in_channels, num_classes, init_features))


# ---------------------------------------------------------------------------------------
def dice_loss(logits, target, smooth=1.0):
    """Restated Dice loss (loss/loss_function.py:74-120): mean over classes of the batch-mean
    1 - (2 sum(p t) + s) / (sum(p^2 + t^2) + s) on softmax probabilities."""
    prob = F.softmax(logits, dim=1)
    onehot = F.one_hot(target.clamp_min(0), num_classes=logits.shape[1])
    onehot = onehot.movedim(-1, 1).to(prob.dtype)
    B, C = prob.shape[0], prob.shape[1]
    p = prob.reshape(B, C, -1)
    t = onehot.reshape(B, C, -1)
    num = 2 * (p * t).sum(-1) + smooth
    den = (p.pow(2) + t.pow(2)).sum(-1) + smooth
    return (1 - num / den).mean(0).sum() / C


def glas_batch(batch, size=256, seed=0, device='cpu'):
    """GlaS-shaped synthetic crops: normalised uniform RGB + a random binary mask."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, size, size, generator=g)
    x = (x - torch.tensor(GLAS_MEAN).view(1, 3, 1, 1)) / torch.tensor(GLAS_STD).view(1, 3, 1, 1)
    m = torch.randint(0, 2, (batch, size, size), generator=g)
    return x.to(device), m.to(device)


def la_batch(batch, shape=(96, 96, 80), seed=0, device='cpu'):
    """LA-shaped synthetic volumes: z-normalised noise + a random binary mask."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 1, *shape, generator=g)
    m = torch.randint(0, 2, (batch, *shape), generator=g)
    return x.to(device), m.to(device)


def deterministic_state_(model, seed=1234):
    """Overwrite every floating-point parameter with values that depend only on its name-order
    index and shape (CPU generator), so two structurally equal nets built in different orders —
    the reference's and ours — get bit-identical weights without shipping a checkpoint.
    BatchNorm scales are centred on 1 and running stats left at their defaults."""
    with torch.no_grad():
        for i, (name, prm) in enumerate(model.named_parameters()):
            g = torch.Generator().manual_seed(seed + i)
            v = torch.randn(prm.shape, generator=g)
            if prm.dim() == 1:
                v = (1.0 + 0.1 * v) if name.endswith('weight') else 0.05 * v
            else:
                fan_in = prm[0].numel()
                v = v * (2.0 / fan_in) ** 0.5
            prm.copy_(v.to(prm.device))
    return model


def disable_dropout_(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    return model
