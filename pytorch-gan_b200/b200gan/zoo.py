"""The model families of the hot path, built from an `nn` namespace (b200gan.nn by default).

These restate the architectures of the reference scripts so that bench.py / smoke() can run on a
box where /root/reference does not exist; with the launcher (launch.py) the reference's own
classes are used unmodified instead.  Citations: implementations/<name>/...
"""
import types

import torch.nn as tnn

from . import nn as bnn


def namespace(stock=False):
    """An object exposing Conv2d, BatchNorm2d, ... : b200gan drop-ins, or stock torch.nn."""
    ns = types.SimpleNamespace()
    for name in dir(tnn):
        setattr(ns, name, getattr(tnn, name))
    if not stock:
        for name, cls in bnn.REPLACEMENTS.items():
            setattr(ns, name, cls)
    return ns


def weights_init_normal(m):
    """dcgan.py:36-42 / pix2pix/models.py:6-12: dispatch on the class *name*."""
    import torch
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm2d") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


class DCGANGenerator(tnn.Module):
    """dcgan/dcgan.py:45-70."""

    def __init__(self, img_size=64, latent_dim=100, channels=1, nn=None):
        super().__init__()
        nn = nn or namespace()
        self.init_size = img_size // 4
        self.l1 = nn.Sequential(nn.Linear(latent_dim, 128 * self.init_size ** 2))
        self.conv_blocks = nn.Sequential(
            nn.BatchNorm2d(128),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 128, 3, stride=1, padding=1),
            nn.BatchNorm2d(128, 0.8),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 64, 3, stride=1, padding=1),
            nn.BatchNorm2d(64, 0.8),
            nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(64, channels, 3, stride=1, padding=1),
            nn.Tanh(),
        )

    def forward(self, z):
        out = self.l1(z)
        out = out.view(out.shape[0], 128, self.init_size, self.init_size)
        return self.conv_blocks(out)


class DCGANDiscriminator(tnn.Module):
    """dcgan/dcgan.py:73-99."""

    def __init__(self, img_size=64, channels=1, nn=None):
        super().__init__()
        nn = nn or namespace()

        def block(cin, cout, bn=True):
            layers = [nn.Conv2d(cin, cout, 3, 2, 1), nn.LeakyReLU(0.2, inplace=True), nn.Dropout2d(0.25)]
            if bn:
                layers.append(nn.BatchNorm2d(cout, 0.8))
            return layers

        self.model = nn.Sequential(*block(channels, 16, bn=False), *block(16, 32), *block(32, 64), *block(64, 128))
        ds_size = img_size // 2 ** 4
        self.adv_layer = nn.Sequential(nn.Linear(128 * ds_size ** 2, 1), nn.Sigmoid())

    def forward(self, img):
        out = self.model(img)
        out = out.view(out.shape[0], -1)
        return self.adv_layer(out)


class WGANGPGenerator(tnn.Module):
    """wgan_gp/wgan_gp.py:42-65 (MLP; BatchNorm1d(out, 0.8): second positional argument is eps)."""

    def __init__(self, img_shape=(1, 32, 32), latent_dim=100, nn=None):
        super().__init__()
        nn = nn or namespace()
        self.img_shape = tuple(img_shape)

        def block(i, o, normalize=True):
            layers = [nn.Linear(i, o)]
            if normalize:
                layers.append(nn.BatchNorm1d(o, 0.8))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            return layers

        n_out = 1
        for v in self.img_shape:
            n_out *= v
        self.model = nn.Sequential(*block(latent_dim, 128, normalize=False), *block(128, 256), *block(256, 512),
                                   *block(512, 1024), nn.Linear(1024, n_out), nn.Tanh())

    def forward(self, z):
        img = self.model(z)
        return img.view(img.shape[0], *self.img_shape)


class WGANGPDiscriminator(tnn.Module):
    """wgan_gp/wgan_gp.py:68-83."""

    def __init__(self, img_shape=(1, 32, 32), nn=None):
        super().__init__()
        nn = nn or namespace()
        n_in = 1
        for v in img_shape:
            n_in *= v
        self.model = nn.Sequential(nn.Linear(n_in, 512), nn.LeakyReLU(0.2, inplace=True), nn.Linear(512, 256),
                                   nn.LeakyReLU(0.2, inplace=True), nn.Linear(256, 1))

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))
