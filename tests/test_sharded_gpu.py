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


def test_files_sharded_over_two_gpus(golden, tmp_path):
    """from_files_to_files(gpu=[0, 1]): the packed int16 corpus is cut into
    contiguous zero-copy shards, one worker thread per GPU; outputs equal the
    single-GPU run bit for bit"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import emphases_b200 as emphases
    emphases.reset_configuration()
    path = tmp_path / 'checkpoint.pt'
    torch.save({'model': state_from_golden(golden('c1'))}, path)
    text_files, audio_files = [], []
    for seed in range(9):
        times, audio = oracle.synthetic_utterance(1600 + seed, duration=1.0 + seed / 2)
        emphases.load.save_wav(tmp_path / f'u{seed}.wav', audio)
        emphases.Alignment.from_times(times).save(tmp_path / f'u{seed}.TextGrid')
        text_files.append(tmp_path / f'u{seed}.TextGrid')
        audio_files.append(tmp_path / f'u{seed}.wav')
    for name, gpu in (('one', 0), ('two', [0, 1])):
        (tmp_path / name).mkdir()
        emphases.from_files_to_files(
            text_files, audio_files, [tmp_path / name / f'u{seed}' for seed in range(9)],
            checkpoint=path, gpu=gpu)
    for seed in range(9):
        a = torch.load(tmp_path / 'one' / f'u{seed}.pt')
        b = torch.load(tmp_path / 'two' / f'u{seed}.pt')
        assert torch.equal(a, b)
