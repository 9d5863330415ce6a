"""Pin the oracle to the LIVE reference at the reference-default size (1.008 B denoiser, f4 VQGAN).
Runs only where /root/reference exists (the dev container); skipped on the GPU box."""
import os
import sys

import pytest
import torch

REF = os.environ.get("PAELLA_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg
    return mg.load_reference()


def test_default_denoiser_forward(ref):
    from oracle import paella_oracle as po
    from paella_b200.synth import rerandomize_, synthetic_conditioning
    ref_modules = ref[0]
    torch.manual_seed(0)
    with torch.no_grad():
        m = ref_modules.Paella(byt5_embd=2560).eval()
        rerandomize_(m.state_dict(), seed=0)
        assert sum(p.numel() for p in m.parameters()) == 1008350592
        cond, _ = synthetic_conditioning(1, 16, with_clip_image=True)
        x = torch.randint(0, 8192, (1, 32, 32), generator=torch.Generator().manual_seed(1))
        r = torch.tensor([0.6])
        want = m(x, r, cond["byt5"], clip=cond["clip"], clip_image=cond["clip_image"])
        got = po.paella_forward(m.state_dict(), po.PaellaConfig(byt5_embd=2560), x, r, cond["byt5"], cond["clip"], cond["clip_image"])
    assert float(want.std()) > 0.05          # re-randomised: logits are not identically zero
    torch.testing.assert_close(got, want, rtol=1e-3, atol=2e-5)


def test_default_vqgan_conv_stacks(ref):
    from oracle import vqgan_oracle as vo
    from paella_b200.synth import rerandomize_
    ref_vqgan = ref[3]
    torch.manual_seed(0)
    with torch.no_grad():
        vq = ref_vqgan.VQModel().eval()
        rerandomize_(vq.state_dict(), seed=4)
        sd = vq.state_dict()
        img = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(7))
        lat = vq.down_blocks(vq.in_block(img))
        torch.testing.assert_close(vo.encode_latents(sd, img).permute(0, 3, 1, 2), lat, rtol=1e-3, atol=2e-5)
        idx = torch.randint(0, 8192, (1, 16, 16), generator=torch.Generator().manual_seed(8))
        torch.testing.assert_close(vo.decode_indices(sd, idx), vq.decode_indices(idx), rtol=1e-3, atol=2e-5)
