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
