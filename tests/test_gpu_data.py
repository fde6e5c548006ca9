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


# ------------------------------------------------------------------ token side (dataset.py:19-31)
def _encoded(N, L, seed):
    """What `tokenizer(list(captions), padding=True, ...)` leaves in `self.encoded_captions`: lists of N rows of one
    length L (ids 101 ... 102 then padding zeros, mask ones then zeros)."""
    rng = np.random.default_rng(seed)
    ids, mask = [], []
    for _ in range(N):
        n = int(rng.integers(min(3, L), L + 1))
        row = ([101] + [int(v) for v in rng.integers(1000, 30000, size=max(n - 2, 0))] + [102])[:n] + [0] * (L - n)
        ids.append(row)
        mask.append([1] * n + [0] * (L - n))
    return {"input_ids": ids, "attention_mask": mask}


@pytest.mark.parametrize("N,L", [(1000, 25), (257, 200), (40, 7), (64, 1)])
def test_token_feed_matches_reference_getitem_and_collate(N, L):
    """TokenFeed.batch == the reference's per-sample torch.tensor(values[idx]) + default collate, bit for bit
    (oracle/datafeed_ref.py restates it; below it is also rebuilt with torch's own default_collate)."""
    import mae_clip_b200.data as d
    from torch.utils.data import default_collate
    enc = _encoded(N, L, N + L)
    feed = d.TokenFeed(enc)
    assert len(feed) == N
    rng = np.random.default_rng(1)
    for idx in ([0], [N - 1, 0, 3 % N, 3 % N], list(rng.integers(0, N, size=300)), [-1, -N], []):
        got = feed.batch(idx)
        ref = datafeed_ref.token_batch_ref(enc, idx)
        for k in ("input_ids", "attention_mask"):
            assert got[k].dtype == torch.int64 and tuple(got[k].shape) == ref[k].shape
            assert np.array_equal(got[k].cpu().numpy(), ref[k])
        if idx:
            items = [{k: torch.tensor(v[i]) for k, v in enc.items()} for i in idx]      # dataset.py:25-28
            col = default_collate(items)
            assert torch.equal(got["input_ids"].cpu(), col["input_ids"]) and torch.equal(got["attention_mask"].cpu(), col["attention_mask"])
    dev_idx = torch.tensor([5 % N, 1 % N], device="cuda")
    assert torch.equal(feed.batch(dev_idx)["input_ids"].cpu(), torch.tensor(enc["input_ids"])[dev_idx.cpu()])
    feed.check()


def test_token_feed_out_of_range_index_raises():
    import mae_clip_b200.data as d
    feed = d.TokenFeed(_encoded(10, 8, 0))
    out = feed.batch([3, 10])
    assert bool((out["input_ids"][1] == 0).all())
    with pytest.raises(IndexError):
        feed.check()
    feed.check()       # the flag was cleared
