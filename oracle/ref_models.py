"""CPU restatement of the reference model definitions and training steps with STOCK torch.nn.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each class cites the reference lines it follows;
equality with the reference's own classes (same seed -> identical parameters, outputs, gradients)
is asserted by oracle/make_golden.py in the build container.
"""
import numpy as np
import torch
import torch.nn as nn


def weights_init_normal(m):
    # dcgan/dcgan.py:36-42, pix2pix/models.py:6-12 (name-based dispatch)
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm2d") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


class DCGANGenerator(nn.Module):
    # dcgan/dcgan.py:45-70
    def __init__(self, img_size=64, latent_dim=100, channels=1):
        super().__init__()
        self.init_size = img_size // 4
        self.l1 = nn.Sequential(nn.Linear(latent_dim, 128 * self.init_size ** 2))
        self.conv_blocks = nn.Sequential(
            nn.BatchNorm2d(128),
            nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 128, 3, stride=1, padding=1),
            nn.BatchNorm2d(128, 0.8),  # second positional argument is eps
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


class DCGANDiscriminator(nn.Module):
    # dcgan/dcgan.py:73-99
    def __init__(self, img_size=64, channels=1):
        super().__init__()

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


def build_dcgan(img_size=64, latent_dim=100, channels=1, seed=0):
    """Construction + init in the reference's order (dcgan.py:106-116) under a fixed torch seed."""
    torch.manual_seed(seed)
    g = DCGANGenerator(img_size, latent_dim, channels)
    d = DCGANDiscriminator(img_size, channels)
    g.apply(weights_init_normal)
    d.apply(weights_init_normal)
    return g, d


def dcgan_step(generator, discriminator, opt_g, opt_d, real_imgs, z, loss=None):
    """One training step exactly as dcgan/dcgan.py:146-183 (labels :147-148, G :157-169, D :175-183).
    Returns (g_loss, d_loss, gen_imgs) as tensors (no host sync)."""
    loss = loss or torch.nn.BCELoss()
    n = real_imgs.shape[0]
    valid = torch.ones(n, 1, device=real_imgs.device)
    fake = torch.zeros(n, 1, device=real_imgs.device)
    opt_g.zero_grad()
    gen_imgs = generator(z)
    g_loss = loss(discriminator(gen_imgs), valid)
    g_loss.backward()
    opt_g.step()
    opt_d.zero_grad()
    real_loss = loss(discriminator(real_imgs), valid)
    fake_loss = loss(discriminator(gen_imgs.detach()), fake)
    d_loss = (real_loss + fake_loss) / 2
    d_loss.backward()
    opt_d.step()
    return g_loss.detach(), d_loss.detach(), gen_imgs.detach()


def make_adam(params, lr=0.0002, b1=0.5, b2=0.999, **kw):
    # dcgan.py:134-135
    return torch.optim.Adam(params, lr=lr, betas=(b1, b2), **kw)


def synthetic_images(n, c, h, w, seed=0):
    """Images in the range of Normalize([0.5],[0.5]) output (dcgan.py:126): uniform [-1, 1)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, c, h, w, generator=g) * 2 - 1


def synthetic_z(n, latent_dim=100, seed=0):
    # dcgan.py:160 draws z with numpy
    rng = np.random.RandomState(seed)
    return torch.tensor(rng.normal(0, 1, (n, latent_dim)), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------
# WGAN-GP (BASELINE config 2)
# ------------------------------------------------------------------------------------------------
class WGANGPGenerator(nn.Module):
    # wgan_gp/wgan_gp.py:42-65
    def __init__(self, img_shape=(1, 32, 32), latent_dim=100):
        super().__init__()
        self.img_shape = tuple(img_shape)

        def block(i, o, normalize=True):
            layers = [nn.Linear(i, o)]
            if normalize:
                layers.append(nn.BatchNorm1d(o, 0.8))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            return layers

        self.model = nn.Sequential(*block(latent_dim, 128, normalize=False), *block(128, 256), *block(256, 512),
                                   *block(512, 1024), nn.Linear(1024, int(np.prod(img_shape))), nn.Tanh())

    def forward(self, z):
        img = self.model(z)
        return img.view(img.shape[0], *self.img_shape)


class WGANGPDiscriminator(nn.Module):
    # wgan_gp/wgan_gp.py:68-83
    def __init__(self, img_shape=(1, 32, 32)):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(int(np.prod(img_shape)), 512), nn.LeakyReLU(0.2, inplace=True),
                                   nn.Linear(512, 256), nn.LeakyReLU(0.2, inplace=True), nn.Linear(256, 1))

    def forward(self, img):
        return self.model(img.view(img.shape[0], -1))


def build_wgan_gp(img_size=32, channels=1, latent_dim=100, seed=0):
    """wgan_gp.py:90-91: G then D, default torch init (no weights_init_normal in this script)."""
    torch.manual_seed(seed)
    shape = (channels, img_size, img_size)
    return WGANGPGenerator(shape, latent_dim), WGANGPDiscriminator(shape)


def compute_gradient_penalty(D, real_samples, fake_samples, alpha):
    """wgan_gp.py:119-138 with the numpy draw of alpha (:122) passed in."""
    interpolates = (alpha * real_samples + ((1 - alpha) * fake_samples)).requires_grad_(True)
    d_interpolates = D(interpolates)
    fake = torch.ones(real_samples.shape[0], 1, device=real_samples.device)
    gradients = torch.autograd.grad(outputs=d_interpolates, inputs=interpolates, grad_outputs=fake,
                                    create_graph=True, retain_graph=True, only_inputs=True)[0]
    gradients = gradients.view(gradients.size(0), -1)
    return ((gradients.norm(2, dim=1) - 1) ** 2).mean()


def synthetic_alpha(n, seed=0):
    rng = np.random.RandomState(seed)
    return torch.tensor(rng.random_sample((n, 1, 1, 1)), dtype=torch.float32)
