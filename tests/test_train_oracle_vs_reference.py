"""Pins oracle/train_oracle.py against the UNMODIFIED reference where it is importable (/root/reference, build container):
`loss_forward` must reproduce train_utils/losses.py:Loss.forward bit for bit on seeded inputs (values and gradients)."""
import pytest
import torch

from oracle import ref_import, train_oracle as TO

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="needs the reference tree (/root/reference)")


@pytest.mark.parametrize("B,M,rows,seed", [(3, 77, 12, 0), (1, 5, 4, 1), (8, 24, 32, 2)])
def test_loss_restatement_matches_reference(B, M, rows, seed):
    ref_import.import_reference()
    from train_utils.losses import Loss               # the reference's own module (train.py:139)
    g = torch.Generator().manual_seed(seed)
    mel_t = (torch.randn(B, 80, M, generator=g) * 2 - 5).clamp_min(-11.5129)
    base = [mel_t + 0.1 * torch.randn(B, 80, M, generator=g), mel_t + 0.2 * torch.randn(B, 80, M, generator=g),
            torch.randn(B, M, 1, generator=g) * 3, torch.softmax(torch.randn(rows, 501, generator=g) * 2, -1)]
    gate_t = torch.zeros(B, M); gate_t[:, -2:] = 1
    a = [t.clone().requires_grad_(True) for t in base]
    b = [t.clone().requires_grad_(True) for t in base]
    out = Loss()([a[0], a[1], a[2], None, torch.zeros(B, M, 29), a[3], torch.full((B,), 29)], (mel_t.clone(), gate_t.clone()))
    mine = TO.loss_forward([b[0], b[1], b[2], None, None, b[3]], (mel_t, gate_t))
    assert set(out) == set(mine) == {"KLD", "mel_loss", "postnet_mel_loss", "gate_loss"}
    for k in out:
        assert torch.equal(out[k], mine[k]), k
    sum(out.values()).backward()                        # train.py:174,184
    sum(mine.values()).backward()
    for x, y in zip(a, b):
        assert torch.equal(x.grad, y.grad)
