"""Run an UNMODIFIED reference script (implementations/<name>/<name>.py) on the b200gan drop-in modules.

    python -m b200gan.launch /path/to/implementations/dcgan/dcgan.py [--b200-iters 4] [--b200-seed 0]
                             [--b200-stock] -- --n_epochs 1 --img_size 64 --batch_size 128

What it does, without touching the script (SURVEY.md sections 0.2, 0.9, 5):
  * seeds the four RNGs the scripts draw from (python `random`, numpy, torch CPU, torch CUDA);
  * rebinds the torch.nn leaf classes to the b200gan drop-ins (skipped with --b200-stock, which runs the
    reference on stock torch for comparison);
  * replaces the network download of `torchvision.datasets.MNIST` (dcgan.py:121, wgan_gp.py:98, gan.py:98) by a
    synthetic dataset of `iters * batch_size` seeded images, and writes synthetic image files for the
    file-based datasets of pix2pix / cyclegan (`../../data/<dataset_name>/{train,test,val}`);
  * runs the script with `runpy` from a scratch working directory two levels deep (the scripts create
    `images/`, `saved_models/` and `../../data/...` relative to the cwd), with the real script directory on
    sys.path so that `from models import *` / `from datasets import *` keep working.
Returns the script's globals (models, optimizers, ...) to the caller of `run()`.
"""
import argparse
import contextlib
import io
import os
import random
import runpy
import sys
import tempfile

import numpy as np
import torch


class SyntheticMNIST(torch.utils.data.Dataset):
    """Same item protocol as torchvision.datasets.MNIST: (PIL 'L' 28x28 image -> transform, int label)."""
    n_items = 64
    seed = 0

    def __init__(self, root=None, train=True, transform=None, target_transform=None, download=False):
        self.transform = transform
        rng = np.random.RandomState(self.seed)
        self.data = rng.randint(0, 256, size=(self.n_items, 28, 28), dtype=np.uint8)
        self.targets = rng.randint(0, 10, size=(self.n_items,))

    def __len__(self):
        return self.n_items

    def __getitem__(self, i):
        from PIL import Image
        img = Image.fromarray(self.data[i], mode="L")
        if self.transform is not None:
            img = self.transform(img)
        return img, int(self.targets[i])


def _write_image_folder(root, layout, n, size, seed):
    """layout 'pix2pix': train/ test/ val/ hold side-by-side A|B images (pix2pix/datasets.py:15-24);
    'cyclegan': train/A train/B test/A test/B (cyclegan/datasets.py:21-22)."""
    from PIL import Image
    rng = np.random.RandomState(seed)
    dirs = ["train", "test", "val"] if layout == "pix2pix" else ["train/A", "train/B", "test/A", "test/B"]
    for d in dirs:
        os.makedirs(os.path.join(root, d), exist_ok=True)
        for i in range(n):
            w = size * 2 if layout == "pix2pix" else size
            arr = rng.randint(0, 256, size=(size, w, 3), dtype=np.uint8)
            Image.fromarray(arr, mode="RGB").save(os.path.join(root, d, f"{i:04d}.png"))


def _opt(argv, name, default):
    if name in argv:
        return type(default)(argv[argv.index(name) + 1])
    return default


def run(script, script_args=(), iters=4, seed=0, stock=False, quiet=False):
    script = os.path.abspath(script)
    sdir, sname = os.path.dirname(script), os.path.basename(os.path.dirname(script))
    script_args = list(script_args)
    batch = _opt(script_args, "--batch_size", {"pix2pix": 1, "cyclegan": 1}.get(sname, 64))

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)

    tmp = tempfile.mkdtemp(prefix="b200gan_launch_")
    work = os.path.join(tmp, "implementations", sname)
    os.makedirs(work)
    if sname in ("pix2pix", "cyclegan"):
        ds_name = _opt(script_args, "--dataset_name", "facades" if sname == "pix2pix" else "monet2photo")
        size = _opt(script_args, "--img_height", 256)
        _write_image_folder(os.path.join(tmp, "data", ds_name), sname, max(iters * batch, 1), size, seed)

    import torchvision.datasets as tvd
    SyntheticMNIST.n_items, SyntheticMNIST.seed = max(iters * batch, 1), seed
    saved = (sys.argv, os.getcwd(), tvd.MNIST, list(sys.path))
    tvd.MNIST = SyntheticMNIST
    sys.argv = [script] + script_args
    sys.path.insert(0, sdir)
    os.chdir(work)
    for m in ("models", "datasets", "utils"):  # the scripts' sibling modules are not packages: no stale copies
        sys.modules.pop(m, None)
    from .patch import patch as _do_patch, unpatch as _do_unpatch
    if not stock:
        _do_patch()
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext():
            g = runpy.run_path(script, run_name="__main__")
    finally:
        if not stock:
            _do_unpatch()
        sys.argv, cwd, tvd.MNIST, sys.path[:] = saved[0], saved[1], saved[2], saved[3]
        os.chdir(cwd)
        for m in ("models", "datasets", "utils"):
            sys.modules.pop(m, None)
    g["__b200_stdout__"] = out.getvalue()
    g["__b200_workdir__"] = work
    return g


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("script")
    ap.add_argument("--b200-iters", type=int, default=4, help="batches of synthetic data per epoch")
    ap.add_argument("--b200-seed", type=int, default=0)
    ap.add_argument("--b200-stock", action="store_true", help="do not patch torch.nn (stock reference run)")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments for the script (after --)")
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    run(a.script, rest, a.b200_iters, a.b200_seed, a.b200_stock)


if __name__ == "__main__":
    main()
