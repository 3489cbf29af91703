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


# ------------------------------------------------------------------------------------------------
# Pix2Pix (BASELINE config 3) and CycleGAN (config 4), stock torch.nn
# ------------------------------------------------------------------------------------------------
def weights_init_normal_cyclegan(m):
    # cyclegan/models.py:6-14 (conv biases zeroed as well)
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
        if hasattr(m, "bias") and m.bias is not None:
            torch.nn.init.constant_(m.bias.data, 0.0)
    elif classname.find("BatchNorm2d") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


class UNetDown(nn.Module):
    # pix2pix/models.py:20-32
    def __init__(self, in_size, out_size, normalize=True, dropout=0.0):
        super().__init__()
        layers = [nn.Conv2d(in_size, out_size, 4, 2, 1, bias=False)]
        if normalize:
            layers.append(nn.InstanceNorm2d(out_size))
        layers.append(nn.LeakyReLU(0.2))
        if dropout:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class UNetUp(nn.Module):
    # pix2pix/models.py:35-52
    def __init__(self, in_size, out_size, dropout=0.0):
        super().__init__()
        layers = [nn.ConvTranspose2d(in_size, out_size, 4, 2, 1, bias=False), nn.InstanceNorm2d(out_size),
                  nn.ReLU(inplace=True)]
        if dropout:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x, skip_input):
        return torch.cat((self.model(x), skip_input), 1)


class GeneratorUNet(nn.Module):
    # pix2pix/models.py:55-101
    def __init__(self, in_channels=3, out_channels=3):
        super().__init__()
        self.down1 = UNetDown(in_channels, 64, normalize=False)
        self.down2 = UNetDown(64, 128)
        self.down3 = UNetDown(128, 256)
        self.down4 = UNetDown(256, 512, dropout=0.5)
        self.down5 = UNetDown(512, 512, dropout=0.5)
        self.down6 = UNetDown(512, 512, dropout=0.5)
        self.down7 = UNetDown(512, 512, dropout=0.5)
        self.down8 = UNetDown(512, 512, normalize=False, dropout=0.5)
        self.up1 = UNetUp(512, 512, dropout=0.5)
        self.up2 = UNetUp(1024, 512, dropout=0.5)
        self.up3 = UNetUp(1024, 512, dropout=0.5)
        self.up4 = UNetUp(1024, 512, dropout=0.5)
        self.up5 = UNetUp(1024, 256)
        self.up6 = UNetUp(512, 128)
        self.up7 = UNetUp(256, 64)
        self.final = nn.Sequential(nn.Upsample(scale_factor=2), nn.ZeroPad2d((1, 0, 1, 0)),
                                   nn.Conv2d(128, out_channels, 4, padding=1), nn.Tanh())

    def forward(self, x):
        d1 = self.down1(x)
        d2 = self.down2(d1)
        d3 = self.down3(d2)
        d4 = self.down4(d3)
        d5 = self.down5(d4)
        d6 = self.down6(d5)
        d7 = self.down7(d6)
        d8 = self.down8(d7)
        u1 = self.up1(d8, d7)
        u2 = self.up2(u1, d6)
        u3 = self.up3(u2, d5)
        u4 = self.up4(u3, d4)
        u5 = self.up5(u4, d3)
        u6 = self.up6(u5, d2)
        u7 = self.up7(u6, d1)
        return self.final(u7)


class Pix2PixDiscriminator(nn.Module):
    # pix2pix/models.py:109-133
    def __init__(self, in_channels=3):
        super().__init__()

        def block(i, o, normalization=True):
            layers = [nn.Conv2d(i, o, 4, stride=2, padding=1)]
            if normalization:
                layers.append(nn.InstanceNorm2d(o))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            return layers

        self.model = nn.Sequential(*block(in_channels * 2, 64, normalization=False), *block(64, 128), *block(128, 256),
                                   *block(256, 512), nn.ZeroPad2d((1, 0, 1, 0)),
                                   nn.Conv2d(512, 1, 4, padding=1, bias=False))

    def forward(self, img_A, img_B):
        return self.model(torch.cat((img_A, img_B), 1))


def build_pix2pix(seed=0):
    """pix2pix.py:60-61,75-76: GeneratorUNet(), Discriminator(), then .apply(weights_init_normal)."""
    torch.manual_seed(seed)
    g, d = GeneratorUNet(), Pix2PixDiscriminator()
    g.apply(weights_init_normal)
    d.apply(weights_init_normal)
    return g, d


class ResidualBlock(nn.Module):
    # cyclegan/models.py:22-37
    def __init__(self, in_features):
        super().__init__()
        self.block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(in_features, in_features, 3),
                                   nn.InstanceNorm2d(in_features), nn.ReLU(inplace=True), nn.ReflectionPad2d(1),
                                   nn.Conv2d(in_features, in_features, 3), nn.InstanceNorm2d(in_features))

    def forward(self, x):
        return x + self.block(x)


class GeneratorResNet(nn.Module):
    # cyclegan/models.py:40-87
    def __init__(self, input_shape, num_residual_blocks):
        super().__init__()
        channels = input_shape[0]
        out_features = 64
        model = [nn.ReflectionPad2d(channels), nn.Conv2d(channels, out_features, 7), nn.InstanceNorm2d(out_features),
                 nn.ReLU(inplace=True)]
        in_features = out_features
        for _ in range(2):
            out_features *= 2
            model += [nn.Conv2d(in_features, out_features, 3, stride=2, padding=1), nn.InstanceNorm2d(out_features),
                      nn.ReLU(inplace=True)]
            in_features = out_features
        for _ in range(num_residual_blocks):
            model += [ResidualBlock(out_features)]
        for _ in range(2):
            out_features //= 2
            model += [nn.Upsample(scale_factor=2), nn.Conv2d(in_features, out_features, 3, stride=1, padding=1),
                      nn.InstanceNorm2d(out_features), nn.ReLU(inplace=True)]
            in_features = out_features
        model += [nn.ReflectionPad2d(channels), nn.Conv2d(out_features, channels, 7), nn.Tanh()]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        return self.model(x)


class CycleGANDiscriminator(nn.Module):
    # cyclegan/models.py:95-122
    def __init__(self, input_shape):
        super().__init__()
        channels, height, width = input_shape
        self.output_shape = (1, height // 2 ** 4, width // 2 ** 4)

        def block(i, o, normalize=True):
            layers = [nn.Conv2d(i, o, 4, stride=2, padding=1)]
            if normalize:
                layers.append(nn.InstanceNorm2d(o))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            return layers

        self.model = nn.Sequential(*block(channels, 64, normalize=False), *block(64, 128), *block(128, 256),
                                   *block(256, 512), nn.ZeroPad2d((1, 0, 1, 0)), nn.Conv2d(512, 1, 4, padding=1))

    def forward(self, img):
        return self.model(img)


def build_cyclegan(input_shape=(3, 64, 64), n_residual_blocks=9, seed=0):
    """cyclegan.py:59-62,80-83: G_AB, G_BA, D_A, D_B constructed in this order, then init in this order."""
    torch.manual_seed(seed)
    g_ab = GeneratorResNet(input_shape, n_residual_blocks)
    g_ba = GeneratorResNet(input_shape, n_residual_blocks)
    d_a = CycleGANDiscriminator(input_shape)
    d_b = CycleGANDiscriminator(input_shape)
    for m in (g_ab, g_ba, d_a, d_b):
        m.apply(weights_init_normal_cyclegan)
    return g_ab, g_ba, d_a, d_b
