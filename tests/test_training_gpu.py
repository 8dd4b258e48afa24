"""Training-step parity: loss and every parameter gradient of the conv model
vs torch autograd through the CPU oracle (fp32)."""
import numpy as np
import pytest
import torch

from golden_util import state_from_golden
from oracle import emphases_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture
def emphases():
    import emphases_b200
    emphases_b200.reset_configuration()
    yield emphases_b200
    emphases_b200.reset_configuration()


def padded_batch(seed=0, items=3):
    generator = torch.Generator().manual_seed(seed)
    lengths = [260, 143, 201][:items]
    words = [9, 4, 6][:items]
    tmax, wmax = max(lengths), max(words)
    features = torch.zeros(items, 80, tmax)
    bounds = torch.zeros(items, 2, wmax, dtype=torch.long)
    for i, (t, w) in enumerate(zip(lengths, words)):
        features[i, :, :t] = torch.randn(80, t, generator=generator)
        cuts = torch.sort(torch.randperm(t - 2, generator=generator)[:w - 1] + 1).values
        edges = torch.cat([torch.tensor([0]), cuts, torch.tensor([t])])
        bounds[i, 0, :w] = edges[:-1]
        bounds[i, 1, :w] = edges[1:]
    targets = torch.rand(items, 1, wmax, generator=generator)
    return (features, torch.tensor(lengths), bounds, torch.tensor(words), targets)


@pytest.mark.parametrize('method', ['sum', 'average', 'max', 'center'])
@pytest.mark.parametrize('loss_fn', ['bce', 'mse'])
def test_gradients_match_autograd(emphases, golden, method, loss_fn):
    if loss_fn == 'mse' and method != 'sum':
        pytest.skip('loss variants are covered with sum pooling')
    emphases.configure(DOWNSAMPLE_METHOD=method, LOSS=loss_fn)
    data = golden('sweep')
    state = state_from_golden(data)
    model = emphases.Model()
    model.load_state_dict({k: v for k, v in state.items() if k in model.state_dict()})
    model = model.cuda().train()
    features, frame_lengths, bounds, word_lengths, targets = padded_batch()

    scores = model(features.cuda(), frame_lengths, bounds, word_lengths)
    assert scores.requires_grad
    value = emphases.loss(
        scores, targets.cuda(), frame_lengths, bounds, word_lengths, training=True)
    value.backward()

    reference = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    expected_scores = oracle.model_forward(
        reference, features, frame_lengths, bounds, word_lengths,
        {'DOWNSAMPLE_METHOD': method})
    expected = oracle.loss(expected_scores, targets, word_lengths, loss_fn)
    expected.backward()

    assert abs(value.item() - expected.item()) < 1e-5 * max(1, abs(expected.item()))
    np.testing.assert_allclose(
        scores.detach().cpu().numpy(), expected_scores.detach().numpy(),
        rtol=1e-5, atol=2e-5)
    for name, parameter in model.named_parameters():
        want = reference[name].grad
        got = parameter.grad.cpu()
        scale = want.abs().max().item() + 1e-12
        error = (got - want).abs().max().item()
        assert error < 2e-4 * scale + 1e-7, (name, error, scale)


def test_train_step_reduces_loss(emphases, golden):
    data = golden('sweep')
    state = state_from_golden(data)
    model = emphases.Model()
    model.load_state_dict({k: v for k, v in state.items() if k in model.state_dict()})
    model = model.cuda()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    batch = padded_batch(1)
    batch = (batch[0].cuda(),) + batch[1:4] + (batch[4].cuda(),)
    losses = [emphases.training.train_step(model, optimizer, batch).item()
              for _ in range(8)]
    assert losses[-1] < losses[0]
