"""Parity at BASELINE.json's full size (config 2: 3000 utterances, 9.2 h,
3.3 M frames) through size-independent properties."""
import numpy as np
import pytest
import torch

from oracle import emphases_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def corpus():
    import bench
    import emphases_b200 as emphases
    from emphases_b200 import engine
    emphases.reset_configuration()
    lengths, times = bench.corpus_layout(3000, seed=1234)
    audio, offsets = bench.make_audio(lengths, seed=99, device='cuda:0')
    plan = engine.make_plan([(t, int(n)) for t, n in zip(times, lengths)], None, 'sum')
    state = bench.random_state()
    model = emphases.Model()
    model.load_state_dict(state)
    model = model.cuda().eval()
    eng = emphases.get_engine(torch.device('cuda', 0))
    return dict(lengths=lengths, times=times, audio=audio, offsets=offsets,
                plan=plan, state=state, weights=model.packed_weights(), eng=eng)


def test_segment_assignment_bit_exact_at_full_size(corpus):
    """Every one of the ~82 k words: pooled frame count and frame-index sum
    (exact small integers in fp32) equal the oracle's [lo, hi)"""
    plan, eng = corpus['plan'], corpus['eng']
    views = eng.upload_plan(plan)
    rows = torch.zeros(plan.total_rows, 80, device='cuda:0')
    starts = torch.from_numpy(plan.row_start.astype(np.int64)).cuda()
    counts = torch.from_numpy(plan.n_rows.astype(np.int64)).cuda()
    owner = torch.repeat_interleave(torch.arange(plan.n_seq, device='cuda:0'), counts)
    within = torch.arange(int(counts.sum()), device='cuda:0') - torch.repeat_interleave(
        torch.cumsum(counts, 0) - counts, counts)
    index = starts[owner] + within
    rows[index, 0] = 1
    rows[index, 1] = within.float()
    pooled = eng.pool(
        rows, views['row_start'], views['n_rows'], views['word_seq'],
        views['word_lo'], views['word_hi'], 'sum').cpu().numpy()
    keep = plan.word_seq >= 0
    lo = plan.word_lo[keep].astype(np.int64)
    hi = np.minimum(plan.word_hi[keep], plan.n_rows[plan.word_seq[keep]]).astype(np.int64)
    assert keep.sum() > 80000
    np.testing.assert_array_equal(pooled[keep, 0].astype(np.int64), hi - lo)
    np.testing.assert_array_equal(
        pooled[keep, 1].astype(np.int64), (lo + hi - 1) * (hi - lo) // 2)
    # and the integer bounds themselves against the oracle on a sample
    for u in range(0, plan.n_seq, 97):
        expected = oracle.word_bounds([tuple(t) for t in corpus['times'][u].tolist()])
        s, n = plan.word_row_start[u], plan.n_words[u]
        assert list(zip(plan.word_lo[s:s + n], plan.word_hi[s:s + n])) == expected


def test_batching_invariance_and_oracle_samples(corpus):
    """The packed 3000-utterance launch equals (i) the same utterances run
    alone and (ii) the CPU oracle, on a sample; fp32 mode, 1e-5"""
    from emphases_b200 import _lib, engine
    plan, eng, weights = corpus['plan'], corpus['eng'], corpus['weights']
    full = eng.forward_packed(corpus['audio'], plan, weights, precision=_lib.PREC_FP32)
    scores = full['scores']
    assert torch.isfinite(scores).all()
    state = corpus['state']
    for u in (0, 1499, 2999):
        offset, count = int(corpus['offsets'][u]), int(corpus['lengths'][u])
        audio = corpus['audio'][offset:offset + count]
        single = engine.make_plan([(corpus['times'][u], count)], None)
        alone = eng.forward_packed(
            audio.clone(), single, weights, precision=_lib.PREC_FP32)['scores']
        s, n = int(plan.word_row_start[u]), int(plan.n_words[u])
        packed = scores[s:s + n]
        assert torch.equal(packed, alone[1:1 + n])      # bit-identical
        expected = oracle.from_alignment_and_audio(
            [tuple(t) for t in corpus['times'][u].tolist()],
            audio.cpu()[None], state)[0]
        assert (packed.cpu() - expected).abs().max() < 1e-5


@pytest.mark.parametrize('precision,tolerance', [
    ('bf16', 2e-3), ('bf16x3', 2e-5), ('bf16x6', 1e-5)])
def test_tensor_core_modes_against_oracle_at_full_size(corpus, precision, tolerance):
    """The benchmarked launch (3000 utterances, one packed batch) in every
    tensor-core mode against the CPU ORACLE on 32 sampled utterances, with
    gain-scaled weights (random-init weights leave every score near 0.5 and
    would hide a broken layer)"""
    import emphases_b200 as emphases
    from emphases_b200 import _lib
    plan, eng = corpus['plan'], corpus['eng']
    torch.manual_seed(7)
    emphases.reset_configuration()
    model = emphases.Model()
    for parameter in model.parameters():
        if parameter.dim() > 1:
            parameter.data.mul_(1.5)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    weights = model.cuda().eval().packed_weights()
    code = {'bf16': _lib.PREC_BF16_TC, 'bf16x3': _lib.PREC_BF16X3_TC,
            'bf16x6': _lib.PREC_BF16X6_TC}[precision]
    scores = eng.forward_packed(corpus['audio'], plan, weights, precision=code)['scores']
    assert torch.isfinite(scores).all()
    keep = torch.from_numpy(plan.word_seq >= 0).cuda()
    assert (scores[~keep] == 0).all()
    spread = scores[keep]
    assert spread.min() < 0.35 and spread.max() > 0.65     # the gain does its job
    key = ('oracle', 7)
    if key not in corpus:
        sample = list(range(0, plan.n_seq, 97)) + [plan.n_seq - 1]
        corpus[key] = sample, [
            oracle.from_alignment_and_audio(
                [tuple(t) for t in corpus['times'][u].tolist()],
                corpus['audio'][int(corpus['offsets'][u]):
                                int(corpus['offsets'][u]) + int(corpus['lengths'][u])
                                ].cpu()[None], state)[0]
            for u in sample]
    sample, expected = corpus[key]
    assert len(sample) >= 30
    worst = 0.
    for u, want in zip(sample, expected):
        s, n = int(plan.word_row_start[u]), int(plan.n_words[u])
        worst = max(worst, (scores[s:s + n].cpu() - want).abs().max().item())
    assert worst < tolerance, (precision, worst)


def test_bf16_mode_tracks_fp32_at_full_size(corpus):
    from emphases_b200 import _lib
    plan, eng, weights = corpus['plan'], corpus['eng'], corpus['weights']
    exact = eng.forward_packed(
        corpus['audio'], plan, weights, precision=_lib.PREC_FP32)['scores']
    fast = eng.forward_packed(
        corpus['audio'], plan, weights, precision=_lib.PREC_BF16_TC)['scores']
    keep = torch.from_numpy(plan.word_seq >= 0).cuda()
    error = (exact - fast).abs()[keep].max().item()
    assert error < 2e-3, error
    assert (exact[~keep] == 0).all() and (fast[~keep] == 0).all()
