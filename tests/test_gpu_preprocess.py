"""CLIP's image preprocessing on the device (ccb_preprocess_image) against the real pipeline the reference uses:
torchvision Compose([Resize(n_px, BICUBIC), CenterCrop(n_px), ToTensor(), Normalize(CLIP mean, std)]) on PIL images
(clip/clip.py `_transform`, applied at inference.py:310; blip_test.py:22-26).  PIL's antialiased bicubic resize works in
22-bit fixed point with uint8 between the passes; the kernels restate that arithmetic, so the result must be bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_transform(n_px):
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode
    return transforms.Compose([
        transforms.Resize(n_px, interpolation=InterpolationMode.BICUBIC),
        transforms.CenterCrop(n_px),
        transforms.ToTensor(),
        transforms.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


@pytest.mark.parametrize("n_px", [224, 64])
def test_preprocess_matches_pil_pipeline(tiny_engine, n_px):
    from PIL import Image
    rng = np.random.default_rng(7)
    sizes = [(480, 640), (640, 480), (224, 224), (225, 300), (1000, 333), (97, 411), (2000, 3000), (n_px, 500), (300, n_px)]
    imgs = []
    for h, w in sizes:
        a = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        a[: h // 2] = (a[: h // 2].astype(np.int32) * np.linspace(0, 1, w)[None, :, None]).astype(np.uint8)   # smooth part
        imgs.append(a)
    tf = reference_transform(n_px)
    want = torch.stack([tf(Image.fromarray(a)) for a in imgs])
    got = tiny_engine.preprocess_images([torch.from_numpy(a) for a in imgs], n_px=n_px).cpu()
    assert got.shape == want.shape == (len(sizes), 3, n_px, n_px)
    for i, (h, w) in enumerate(sizes):
        assert torch.equal(got[i], want[i]), (h, w, (got[i] - want[i]).abs().max().item())


def test_resize_geometry_follows_torchvision(tiny_engine):
    from torchvision.transforms import functional as F
    from PIL import Image
    for h, w in [(480, 640), (640, 480), (225, 300), (97, 411), (1001, 333), (224, 224)]:
        im = Image.fromarray(np.zeros((h, w, 3), dtype=np.uint8))
        r = F.resize(im, 224, interpolation=F.InterpolationMode.BICUBIC)
        nh, nw, top, left = tiny_engine.resize_geometry(h, w, 224)
        assert (nw, nh) == r.size
        c = F.center_crop(r, 224)
        assert c.size == (224, 224)
