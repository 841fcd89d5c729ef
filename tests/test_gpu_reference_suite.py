"""GPU mirrors of the reference's own component tests (/root/reference/tests, each test cites the one it restates): the
same constructions and assertions, tensors on the GPU, through our mirrors of the classes — what a reference user's own
test-suite would exercise after switching packages.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _bundle(n=10, near=2.0, far=4.0):
    from nerfstudio_b200.cameras.rays import RayBundle

    o = torch.zeros(n, 3, device="cuda")
    return RayBundle(origins=o, directions=torch.ones_like(o), pixel_area=torch.ones(n, 1, device="cuda"),
                     nears=torch.full((n, 1), near, device="cuda"), fars=torch.full((n, 1), far, device="cuda"))


@pytest.mark.parametrize("cls_name", ["UniformSampler", "LinearDisparitySampler", "SqrtSampler", "LogSampler"])
def test_spaced_samplers(cuda, cls_name):
    """tests/model_components/test_ray_sampler.py:18-86."""
    from nerfstudio_b200.model_components import ray_samplers

    sampler = getattr(ray_samplers, cls_name)(num_samples=15)
    rs = sampler(_bundle())
    assert rs.frustums.get_positions().shape[-2] == 15


def test_pdf_sampler(cuda):
    """tests/model_components/test_ray_sampler.py:88-112 (does not crash; 15 new samples per ray)."""
    from nerfstudio_b200.model_components.ray_samplers import PDFSampler, UniformSampler

    rb = _bundle()
    coarse = UniformSampler(num_samples=15)(rb)
    weights = torch.ones(10, 15, 1, device="cuda")
    rs = PDFSampler(15)(rb, coarse, weights, 15)  # include_original (default): 16 old + 16 new edges -> 31 samples
    assert rs.frustums.get_positions().shape[-2] == 31
    rs = PDFSampler(15, include_original=False)(rb, coarse, weights, 15)
    assert rs.frustums.get_positions().shape[-2] == 15


def test_nerf_encoder(cuda):
    """tests/field_components/test_encodings.py:29-55."""
    from nerfstudio_b200.field_components.encodings import NeRFEncoding

    enc = NeRFEncoding(in_dim=4, num_frequencies=3, min_freq_exp=0, max_freq_exp=3)
    assert enc.get_out_dim() == 24
    out = enc(torch.ones(2, 3, 4, device="cuda"))
    assert out.shape[-1] == 24 and float(out.max()) == pytest.approx(1.0, abs=1e-6)
    out = enc(torch.zeros(2, 3, 4, device="cuda"))
    assert out.shape[-1] == 24 and float(out.min()) == pytest.approx(0.0, abs=1e-6)


def test_sh_encoder(cuda):
    """tests/field_components/test_encodings.py:125-140."""
    from nerfstudio_b200.field_components.encodings import SHEncoding

    with pytest.raises(ValueError):
        SHEncoding(levels=6)
    enc = SHEncoding(levels=5)
    assert enc.get_out_dim() == 25
    x = torch.zeros(10, 3, device="cuda")
    x[..., 1] = 1
    assert enc(x).shape == (10, 25)


def test_hash_encoder_four_features(cuda):
    """tests/field_components/test_encodings.py:143-169 (F = 4 per level, T = 2^5; torch and tcnn implementations)."""
    from nerfstudio_b200.field_components.encodings import HashEncoding

    x = torch.rand(10, 3, device="cuda")
    for impl in ("torch", "tcnn"):
        enc = HashEncoding(num_levels=4, features_per_level=4, log2_hashmap_size=5, implementation=impl).cuda()
        assert enc.get_out_dim() == 16
        assert enc(x).shape == (10, 16)


def test_mlp(cuda):
    """tests/field_components/test_mlp.py:11-30."""
    from torch import nn

    from nerfstudio_b200.field_components.mlp import MLP

    mlp = MLP(in_dim=6, out_dim=10, num_layers=2, layer_width=32, out_activation=nn.ReLU()).cuda()
    assert mlp.get_out_dim() == 10
    assert mlp(torch.ones(9, 6, device="cuda")).shape[-1] == 10


def test_renderers(cuda):
    """tests/model_components/test_renderers.py:12-26,47-80."""
    from nerfstudio_b200.model_components import renderers

    w = torch.ones(3, 10, 1, device="cuda") / 10
    rgb_r = renderers.RGBRenderer()
    assert float(rgb_r(rgb=torch.ones(3, 10, 3, device="cuda"), weights=w).max()) > 0.9
    assert float(rgb_r(rgb=torch.zeros(3, 10, 3, device="cuda"), weights=w).max()) == pytest.approx(0, abs=1e-6)
    assert float(renderers.AccumulationRenderer()(weights=w).max()) > 0.9


def test_frustum_get_positions(cuda):
    """tests/cameras/test_rays.py:11-30 (known answer)."""
    from nerfstudio_b200.cameras.rays import Frustums

    o = torch.ones(3, device="cuda")[None]
    d = torch.tensor([1.0, 0, 0], device="cuda")[None]
    fr = Frustums(origins=o, directions=d, starts=torch.ones(1, 1, device="cuda"), ends=torch.ones(1, 1, device="cuda") * 2,
                  pixel_area=torch.ones(1, 1, device="cuda"))
    assert fr.get_positions().cpu().flatten().tolist() == pytest.approx([2.5, 1.0, 1.0])
