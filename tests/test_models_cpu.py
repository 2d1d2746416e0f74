"""The two readings of the reference's actor-critic .cfg as plain torch modules: parameter counts
and output shapes.  No CUDA needed: these classes hold no kernels (the tcgen05 variants are covered by the gpu tests)."""
import torch

from xagents_b200.agents.models import AsCodedConv1dCNN, NatureCNN


def test_documented_conv2d_network_has_the_readme_parameter_count():
    net = NatureCNN(4, 6)
    assert sum(p.numel() for p in net.parameters()) == 1_687_719          # README.md:243-259 summary; bench.py's all-reduce size
    actor, critic = net(torch.rand(3, 84, 84, 4))
    assert actor.shape == (3, 6) and critic.shape == (3, 1)


def test_as_coded_conv1d_network_matches_what_model_reader_builds():
    net = AsCodedConv1dCNN(4, 6)
    # Conv1D 32x8/4 -> 64x4/2 -> 64x3/1 along the width of [n, 84, 84, 4] (rows = extended batch), flatten 84*7*64, FC512, heads
    assert sum(p.numel() for p in net.parameters()) == 19_293_351
    x = torch.rand(2, 84, 84, 4)
    actor, critic = net(x)
    assert actor.shape == (2, 6) and critic.shape == (2, 1)
    # rows are independent until the flatten: permuting image rows permutes the trunk's output rows the same way
    perm = torch.randperm(84)
    a = net.trunk(x.permute(0, 3, 1, 2))
    b = net.trunk(x[:, perm].permute(0, 3, 1, 2))
    assert torch.allclose(a[:, :, perm], b, atol=1e-6)


def _torch_gather_cast(src, index_map, out, stream=None):
    idx = index_map.long()
    vals = torch.where(idx >= 0, src[idx.clamp(min=0)], torch.zeros((), dtype=src.dtype))
    out.copy_(vals.to(out.dtype))
    return out


def test_operand_index_map_reproduces_the_relayout_code(monkeypatch):
    """tc_operands.OperandPack: the index map built by running the re-layout code on parameter indices must give, through
    one gather+cast, exactly what the re-layout code gives on the weights -- free-standing parameters and parameters
    re-seated as views of one flat buffer (TorchModel), including a refresh after the weights changed."""
    from xagents_b200 import ops
    from xagents_b200.agents import tc_operands
    monkeypatch.setattr(ops, 'gather_cast_f32', _torch_gather_cast)
    torch.manual_seed(0)
    net = NatureCNN(4, 6)
    for q in net.parameters():
        torch.nn.init.normal_(q)                                      # biases too

    def expected():
        named = dict(c1w=net.trunk[0].weight, c1b=net.trunk[0].bias, c2w=net.trunk[2].weight, c2b=net.trunk[2].bias,
                     c3w=net.trunk[4].weight, c3b=net.trunk[4].bias, fcw=net.trunk[7].weight, fcb=net.trunk[7].bias,
                     aw=net.actor.weight, ab=net.actor.bias, cw=net.critic.weight, cb=net.critic.bias)
        b16, f32 = tc_operands.derive({k: v.detach() for k, v in named.items()}, 0.0)
        return {**{k: v.to(torch.bfloat16) for k, v in b16.items()}, **f32}

    def check(pack):
        for name, want in expected().items():
            got = getattr(pack, name)
            assert got.shape == want.shape and torch.equal(got, want), name

    pack = tc_operands.OperandPack(net)
    check(pack)
    addresses = {name: getattr(pack, name).data_ptr() for name in ('w1', 'wf_t', 'bh')}
    # re-seat the parameters in one flat buffer (what TorchModel does) and change them
    params = list(net.parameters())
    flat = torch.zeros(sum(q.numel() for q in params) + 5)
    off = 3                                                            # not at the start of the storage
    for q in params:
        flat[off:off + q.numel()].copy_(q.data.reshape(-1))
        q.data = flat[off:off + q.numel()].view_as(q)
        off += q.numel()
    flat.mul_(-0.5)
    pack.refresh()
    check(pack)
    assert addresses == {name: getattr(pack, name).data_ptr() for name in addresses}      # captured graphs stay valid
    flat.add_(1.0)
    pack.refresh()
    check(pack)
