"""Image side of the data feed ("next" row 4): uint8 HWC -> normalised fp32 NCHW through the C ABI, bit-exact
against oracle/datafeed_ref.py (albumentations' normalize restated; dataset.py:33-34,49)."""
import numpy as np
import pytest
import torch

from oracle import datafeed_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(4, 224, 224, 3), (1, 7, 5, 3), (3, 32, 30, 3), (0, 8, 8, 3)])
def test_normalize_images_bit_exact(shape):
    import mae_clip_b200.data as d
    rng = np.random.default_rng(sum(shape))
    pix = rng.integers(0, 256, size=shape, dtype=np.uint8)
    out = d.normalize_images(torch.from_numpy(pix).cuda())
    ref = datafeed_ref.normalize_ref(pix)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), ref)       # same two fp32 operations per element


def test_normalize_every_byte_value_and_custom_stats():
    import mae_clip_b200.data as d
    pix = np.arange(256, dtype=np.uint8).repeat(3).reshape(1, 16, 16, 3)
    out = d.normalize_images(torch.from_numpy(pix).cuda(), mean=(0.5, 0.25, 0.0), std=(0.5, 1.0, 2.0), max_pixel_value=255.0)
    assert np.array_equal(out.cpu().numpy(), datafeed_ref.normalize_ref(pix, (0.5, 0.25, 0.0), (0.5, 1.0, 2.0)))
    single = d.normalize_images(torch.from_numpy(pix[0]).cuda())
    assert single.shape == (3, 16, 16)


def test_synthetic_batch_feeds_the_model():
    import mae_clip_b200 as m
    import mae_clip_b200.data as d
    b = d.synthetic_batch(8, size=32, seq_len=25)
    assert b["image"].shape == (8, 3, 32, 32) and b["input_ids"].shape == (8, 25) and b["attention_mask"].dtype == torch.long

    class Tower(torch.nn.Module):
        def __init__(self, n_in, out):
            super().__init__()
            self.lin = torch.nn.Linear(n_in, out)

        def forward(self, x=None, input_ids=None, attention_mask=None):
            x = x if x is not None else input_ids.float()
            return self.lin(x.reshape(x.shape[0], -1))

    model = m.CLIPModel(image_embedding=64, text_embedding=32, image_encoder=Tower(3 * 32 * 32, 64),
                        text_encoder=Tower(25, 32)).cuda()
    loss = model(b)
    loss.backward()
    assert torch.isfinite(loss) and model.image_projection.projection.weight.grad is not None


def test_rejects_wrong_dtype():
    import mae_clip_b200.data as d
    with pytest.raises(ValueError):
        d.normalize_images(torch.zeros(1, 4, 4, 3, device="cuda"))
