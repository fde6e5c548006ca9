"""Inference-side retrieval ("next" row 3): fused normalise + similarity + exact top-k through the C ABI
against oracle/retrieval_ref.py (inference.py:42-46 restated) and against the golden fixture produced by
the reference's own expressions."""
import numpy as np
import pytest
import torch

from oracle import retrieval_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("Q,N,D,k", [(1, 4000, 256, 45), (3, 1000, 256, 1), (40, 777, 256, 10), (2, 64, 128, 64),
                                     (5, 300, 512, 7), (2, 500, 96, 9), (1, 200000, 256, 1024)])
def test_similarity_topk_vs_oracle(Q, N, D, k):
    import mae_clip_b200 as m
    g = torch.Generator().manual_seed(Q * 1000 + N)
    t = torch.randn(Q, D, generator=g) * 3
    x = torch.randn(N, D, generator=g) * torch.rand(N, 1, generator=g) * 5
    ref_vals, ref_idx = retrieval_ref.topk_ref(t, x, k)
    ref_scores = retrieval_ref.similarity_ref(t, x)
    vals, idx, scores = m.similarity_topk(t.cuda(), x.cuda(), k, return_scores=True)
    assert vals.shape == (Q, k) and idx.dtype == torch.int64
    np.testing.assert_allclose(scores.cpu().numpy(), ref_scores.numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(vals.cpu().numpy(), ref_vals.numpy(), rtol=0, atol=2e-6)
    # our own scores: descending, consistent with the indices, and exactly the k largest
    sc = scores.cpu()
    assert torch.equal(torch.gather(sc, 1, idx.cpu()), vals.cpu())
    assert (vals[:, 1:] <= vals[:, :-1]).all()
    kth = vals[:, -1:].cpu()
    assert ((sc > kth).sum(1) < k).all() and ((sc >= kth).sum(1) >= k).all()
    # indices agree with the oracle wherever a reference score is separated from both neighbours by more than rounding
    srt = torch.sort(ref_scores, dim=1, descending=True).values[:, :min(k + 1, N)]
    sep_next = (srt[:, :-1] - srt[:, 1:]) > 1e-5        # position j vs j + 1
    ok = torch.ones(Q, k, dtype=torch.bool)
    ok[:, :sep_next.shape[1]] &= sep_next[:, :k]
    ok[:, 1:] &= sep_next[:, :k - 1]
    assert ok.float().mean() > 0.5 or N > 100000
    assert torch.equal(idx.cpu()[ok], ref_idx[ok])


def test_ties_resolve_to_lower_index_and_nan_first():
    import mae_clip_b200 as m
    x = torch.zeros(10, 128)
    x[:, 0] = 1.0               # all images identical -> all scores equal
    t = torch.zeros(1, 128)
    t[0, 0] = 2.0
    vals, idx = m.similarity_topk(t.cuda(), x.cuda(), 4)
    assert idx.cpu().tolist() == [[0, 1, 2, 3]] and torch.allclose(vals.cpu(), torch.ones(1, 4))
    x[7, 0] = float("nan")
    vals, idx = m.similarity_topk(t.cuda(), x.cuda(), 2)
    assert idx[0, 0].item() == 7 and torch.isnan(vals[0, 0])   # torch.topk ranks NaN highest


def test_zero_rows_use_normalize_eps():
    """F.normalize divides by max(norm, 1e-12): an all-zero embedding scores 0, not NaN."""
    import mae_clip_b200 as m
    x = torch.randn(6, 256)
    x[2] = 0
    t = torch.randn(2, 256)
    _, _, scores = m.similarity_topk(t.cuda(), x.cuda(), 1, return_scores=True)
    ref = retrieval_ref.similarity_ref(t, x)
    assert scores[:, 2].abs().max().item() == 0.0 and torch.allclose(scores.cpu(), ref, atol=2e-6)


def test_find_matches_matches_reference_expression():
    import mae_clip_b200 as m
    from mae_clip_b200.inference import find_matches

    class Tower(torch.nn.Module):
        def __init__(self, out):
            super().__init__()
            self.lin = torch.nn.Linear(16, out)

        def forward(self, x=None, input_ids=None, attention_mask=None):
            return self.lin((x if x is not None else input_ids).float())

    torch.manual_seed(0)
    model = m.CLIPModel(image_embedding=64, text_embedding=32, image_encoder=Tower(64), text_encoder=Tower(32)).cuda().eval()
    bank = torch.randn(500, 256, device="cuda")
    names = [f"img_{i}.jpg" for i in range(500)]
    q = {"input_ids": torch.randint(0, 9, (1, 16)).tolist(), "attention_mask": torch.ones(1, 16, dtype=torch.long).tolist()}
    got = find_matches(model, bank, q, names, n=9)
    with torch.no_grad():
        te = model.text_projection(model.text_encoder(input_ids=torch.tensor(q["input_ids"]).cuda(),
                                                      attention_mask=torch.tensor(q["attention_mask"]).cuda()))
    assert got == retrieval_ref.find_matches_ref(te.cpu(), bank.cpu(), names, n=9)


def test_k_larger_than_bank_raises():
    import mae_clip_b200 as m
    with pytest.raises(RuntimeError):
        m.similarity_topk(torch.randn(1, 256).cuda(), torch.randn(5, 256).cuda(), 6)
