"""Seeded differential test: random utterances, chunk sizes, pooling methods,
downsample locations and input forms through the public API against the CPU
oracle, in EVERY precision mode (north_star tolerances: 1e-5 for the fp32-grade
modes 'fp32' and 'bf16x6', 2e-3 for the 'bf16' tensor-core mode; 'bf16x3' is
held to 2e-5).  Integer segmentation is implied bit-exact: a wrong bound moves
a score by far more than the tolerance.  Weights are gain-scaled (x1.4) so the
scores spread over (0, 1) instead of hugging 0.5."""
import numpy as np
import pytest
import torch

from oracle import emphases_oracle as oracle

pytestmark = pytest.mark.gpu


def _utterance(generator, index):
    duration = float(generator.uniform(0.3, 9.0))
    samples = int(duration * 16000) + int(generator.integers(0, 160))
    words = int(generator.integers(1, max(2, int(3.5 * duration))))
    # cuts on a 2-frame grid keep every word non-empty (empty words are the
    # adversarial pooling tests' business: `max` raises, `average` is NaN)
    grid = np.arange(2, samples // 160 - 1, 2)
    words = min(words, len(grid) + 1)
    cuts = np.sort(generator.choice(grid, size=words - 1, replace=False)) \
        if words > 1 else np.zeros(0)
    edges = np.concatenate([[0.], (cuts + generator.uniform(0.05, 0.9, len(cuts))) * 0.01,
                            [samples / 16000.]])
    if generator.random() < 0.3 and words > 1:
        edges[0] = float(generator.uniform(0.0, edges[1] * 0.5))      # leading gap
    torch_generator = torch.Generator().manual_seed(1000 + index)
    audio = (0.1 * torch.randn(1, samples, generator=torch_generator)).clamp(-1, 1)
    if generator.random() < 0.5:                                      # 16-bit PCM values
        audio = (audio * 32768.).round().clamp(-32768, 32767) / 32768.
    return [tuple(pair) for pair in zip(edges[:-1], edges[1:])], audio


TOLERANCE = {'fp32': 1e-5, 'bf16x6': 1e-5, 'bf16x3': 2e-5, 'bf16': 2e-3}
_expected = {}      # (seed, batch_size) -> oracle scores: shared by the precision modes


@pytest.mark.parametrize('precision', ['fp32', 'bf16x6', 'bf16x3', 'bf16'])
@pytest.mark.parametrize('seed', list(range(8)))
def test_random_corpora_against_oracle(seed, precision):
    import emphases_b200 as emphases
    from emphases_b200 import scheduler
    generator = np.random.default_rng(seed)
    emphases.reset_configuration()
    method = ['sum', 'average', 'max', 'center'][int(generator.integers(0, 4))]
    location = ['intermediate', 'loss', 'inference', 'intermediate'][int(generator.integers(0, 4))]
    emphases.configure(
        DOWNSAMPLE_METHOD=method, DOWNSAMPLE_LOCATION=location, PRECISION=precision,
        MAX_ROWS_PER_LAUNCH=int(generator.choice([700, 2500, 1 << 19])))
    torch.manual_seed(seed)
    model = emphases.Model()
    for parameter in model.parameters():
        if parameter.dim() > 1:
            parameter.data.mul_(1.4)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    config = {'DOWNSAMPLE_METHOD': method, 'DOWNSAMPLE_LOCATION': location}

    items = [_utterance(generator, 100 * seed + i) for i in range(24)]
    alignments = [emphases.Alignment.from_times(times) for times, _ in items]
    audios = [audio for _, audio in items]
    for batch_size in (None, int(generator.integers(40, 400))):
        if (seed, batch_size) not in _expected:
            _expected[seed, batch_size] = [
                oracle.from_alignment_and_audio(times, audio, state, config, batch_size)
                for times, audio in items]
        expected = _expected[seed, batch_size]
        forms = {
            'list': audios,
            'packed fp32': scheduler.pack_audio(audios),
            'single calls': None}
        for name, form in forms.items():
            if form is None:
                got = [
                    emphases.from_alignments_and_audio(
                        [alignment], [audio], 16000, model=model, gpu=0,
                        batch_size=batch_size)[0]
                    for alignment, audio in zip(alignments[:6], audios[:6])]
            else:
                got = emphases.from_alignments_and_audio(
                    alignments, form, 16000, model=model, gpu=0, batch_size=batch_size)
            for index, (result, want) in enumerate(zip(got, expected)):
                assert result.shape == want.shape, (name, index)
                error = (result.cpu() - want).abs().max().item() if want.numel() else 0.
                assert error < TOLERANCE[precision], (
                    seed, precision, method, location, batch_size, name, index, error)
