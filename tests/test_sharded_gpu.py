"""In-process multi-GPU sharding (gpu=[0, 1]): needs >= 2 GPUs"""
import pytest
import torch

from golden_util import state_from_golden
from oracle import emphases_oracle as oracle

pytestmark = pytest.mark.gpu


def test_run_sharded_matches_single_gpu(golden, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import emphases_b200 as emphases
    emphases.reset_configuration()
    path = tmp_path / 'checkpoint.pt'
    torch.save({'model': state_from_golden(golden('c1'))}, path)
    alignments, audios = [], []
    for seed in range(10):
        times, audio = oracle.synthetic_utterance(1500 + seed)
        alignments.append(emphases.Alignment.from_times(times))
        audios.append(audio)
    single = emphases.from_alignments_and_audio(
        alignments, audios, 16000, path, gpu=0)
    sharded = emphases.from_alignments_and_audio(
        alignments, audios, 16000, path, gpu=[0, 1])
    assert len(single) == len(sharded) == 10
    for a, b in zip(single, sharded):
        assert torch.equal(a, b)
