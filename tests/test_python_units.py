"""Pure-Python / CPU unit tests: model definitions, op references, trainer bookkeeping."""
import pytest
import torch

from pslite_b200.models.llama import Llama, LlamaConfig, apply_rope, precompute_rope
from pslite_b200.models.resnet import resnet50, resnet_tiny
from pslite_b200.ops.fused import rope_split_reference, swiglu_reference
from pslite_b200.parallel.ps_trainer import symmetric_layout


def test_llama3_8b_parameter_count():
    cfg = LlamaConfig.llama3_8b()
    assert cfg.num_params() == 8_030_261_248
    with torch.device("meta"):
        m = Llama(cfg)
    assert sum(p.numel() for p in m.parameters()) == cfg.num_params()


def test_tiny_llama_trains_on_cpu():
    torch.manual_seed(0)
    cfg = LlamaConfig.tiny()
    m = Llama(cfg)
    m.init_weights(seed=0)
    opt = torch.optim.AdamW(m.parameters(), lr=3e-3)
    tok = torch.randint(0, cfg.vocab_size, (2, 33))
    losses = []
    for _ in range(5):
        loss = m(tok[:, :-1], tok[:, 1:])
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


def test_checkpointed_and_plain_forward_agree():
    torch.manual_seed(0)
    a = Llama(LlamaConfig.tiny(ckpt_layers=0))
    b = Llama(LlamaConfig.tiny(ckpt_layers=2))
    b.load_state_dict(a.state_dict())
    tok = torch.randint(0, 512, (2, 17))
    la = a(tok[:, :-1], tok[:, 1:])
    lb = b(tok[:, :-1], tok[:, 1:])
    la.backward()
    lb.backward()
    assert torch.allclose(la, lb, atol=1e-6)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa.grad, pb.grad, atol=1e-5)


def test_rope_split_reference_matches_model_rope():
    B, S, H, KV, D = 2, 9, 4, 2, 16
    cos, sin = precompute_rope(D, 16, 10000.0, "cpu")
    qkv = torch.randn(B, S, (H + 2 * KV) * D)
    q, k, v = rope_split_reference(qkv, cos, sin, H, KV, D)
    q0, k0, v0 = qkv.split([H * D, KV * D, KV * D], dim=-1)
    assert torch.allclose(q, apply_rope(q0.view(B, S, H, D), cos[:S], sin[:S]), atol=1e-6)
    assert torch.allclose(k, apply_rope(k0.view(B, S, KV, D), cos[:S], sin[:S]), atol=1e-6)
    assert torch.equal(v, v0.view(B, S, KV, D))
    # rotation preserves norms
    assert torch.allclose(q.norm(dim=-1), q0.view(B, S, H, D).norm(dim=-1), atol=1e-4)


def test_swiglu_reference():
    gu = torch.randn(4, 10)
    g, u = gu.chunk(2, -1)
    assert torch.allclose(swiglu_reference(gu), torch.nn.functional.silu(g) * u, atol=1e-6)


def test_resnet_shapes_and_size():
    m = resnet_tiny()
    assert m(torch.randn(2, 3, 32, 32)).shape == (2, 10)
    with torch.device("meta"):
        big = resnet50()
    n = sum(p.numel() for p in big.parameters())
    assert 25_000_000 < n < 26_000_000


def test_symmetric_layout_alignment():
    params = [torch.empty(5), torch.empty(64), torch.empty(3, 100)]
    offs, total = symmetric_layout(params)
    assert offs == [0, 64, 128] and total == 128 + 320
    assert all(o % 64 == 0 for o in offs)  # 128-byte aligned bf16 offsets


def test_trainer_flat_view_of_channels_last_parameters():
    """the PS moves a parameter as the flat buffer of its storage: for a channels-last convolution weight that
    is a permuted order, and a gradient that arrives in another layout must be brought into the same one"""
    from pslite_b200.parallel.ps_trainer import _flat

    w = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2).contiguous(memory_format=torch.channels_last)
    f = _flat(w)
    assert f.numel() == w.numel() and f.data_ptr() == w.data_ptr() and f.is_contiguous()
    assert torch.equal(f, w.permute(0, 2, 3, 1).reshape(-1))          # storage order: N, H, W, C
    g = torch.randn(2, 3, 2, 2)                                        # a gradient in the default layout
    g2 = torch.empty_like(w).copy_(g)
    assert g2.stride() == w.stride()
    _flat(w).copy_(_flat(g2))                                          # what a pull does with the server's answer
    assert torch.equal(w, g)
    c = torch.randn(4, 5)
    assert _flat(c).data_ptr() == c.data_ptr() and _flat(c).shape == (20,)
    with pytest.raises(AssertionError):
        _flat(torch.randn(4, 6)[:, ::2])                               # not dense: cannot be a PS parameter
