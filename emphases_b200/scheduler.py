"""Corpus scheduling: launch bucketing on one GPU, length-balanced sharding
across GPUs.

Utterances are independent (emphases/core.py:169-179 loops over files), so a
corpus shards by utterance with no collective: each GPU runs its shard
through the packed kernels and the host gathers the per-word scores.
"""
import threading
from typing import List, Sequence

import numpy as np
import torch

import emphases_b200 as emphases
from . import _lib, engine
from .alignment import as_times


###############################################################################
# Host placement
###############################################################################


def bind_to_gpu_numa_node(index):
    """Pin the calling thread to the CPUs local to GPU `index` so that pinned
    staging buffers (first-touch) and the copy-issuing thread sit on the
    GPU's NUMA node; host->device bandwidth halves from the wrong socket.
    Returns the cpulist string, or None when it cannot be determined."""
    import os
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        domain = torch.cuda.get_device_properties(index).pci_domain_id
        device = torch.cuda.get_device_properties(index).pci_device_id
        path = f'/sys/bus/pci/devices/{domain:04x}:{bus:02x}:{device:02x}.0/local_cpulist'
        with open(path) as stream:
            cpulist = stream.read().strip()
        cpus = set()
        for part in cpulist.split(','):
            if '-' in part:
                lo, hi = part.split('-')
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return cpulist
    except (OSError, AttributeError, ValueError, RuntimeError):
        return None


###############################################################################
# Packed host audio
###############################################################################


class PackedAudio:
    """A corpus of mono 16 kHz utterances in ONE pinned host buffer.

    Utterance i occupies samples [offsets[i], offsets[i] + lengths[i]);
    offsets are multiples of engine.AUDIO_ALIGN samples (the kernels' bulk-copy
    and vector-load alignment).
    A launch covering utterances [a, b) needs a single host->device copy.
    """

    def __init__(self, buffer, offsets, lengths):
        self.buffer = buffer
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.lengths = np.asarray(lengths, dtype=np.int64)
        # optional callable(j): blocks until utterances 0..j are resident in
        # `buffer` (a corpus still being decoded in the background)
        self.ready = None

    def __len__(self):
        return len(self.lengths)

    def __getitem__(self, index):
        if self.ready is not None:
            self.ready(int(index))
        start = int(self.offsets[index])
        return self.buffer[start:start + int(self.lengths[index])][None]

    @staticmethod
    def layout(lengths):
        lengths = np.asarray(lengths, dtype=np.int64)
        padded = engine.align_samples(lengths)
        offsets = np.concatenate([[0], np.cumsum(padded[:-1])]) \
            if len(lengths) else np.zeros(0, dtype=np.int64)
        total = int(padded.sum())
        return offsets.astype(np.int64), max(total, engine.AUDIO_ALIGN)


def pack_audio(audios, dtype=torch.float32, pin=True):
    """List of (C, T) tensors -> PackedAudio (channel 0, as mels.py:48)"""
    lengths = [int(audio.shape[-1]) for audio in audios]
    offsets, total = PackedAudio.layout(lengths)
    buffer = torch.zeros(
        total, dtype=dtype, pin_memory=pin and torch.cuda.is_available())
    for offset, audio in zip(offsets, audios):
        buffer[offset:offset + audio.shape[-1]] = audio[0]
    return PackedAudio(buffer, offsets, lengths)


###############################################################################
# Length-balanced scheduling
###############################################################################


def lpt_assign(costs: Sequence[float], workers: int) -> List[List[int]]:
    """Longest-processing-time-first greedy assignment: sort by cost
    descending, always give the next item to the least-loaded worker.
    Returns, per worker, the item indices in their original order."""
    import heapq
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, worker) for worker in range(workers)]
    shards = [[] for _ in range(workers)]
    for index in order:
        load, worker = heapq.heappop(heap)
        shards[worker].append(index)
        heapq.heappush(heap, (load + costs[index], worker))
    return [sorted(shard) for shard in shards]


def bucket_launches(frames: Sequence[int], max_rows: int) -> List[List[int]]:
    """Split utterances (kept in order) into launches of <= max_rows packed
    rows; an utterance longer than max_rows gets a launch of its own."""
    gap = engine.separator_rows()
    need = np.asarray(frames, dtype=np.int64) + gap
    # rows used before utterance i if everything were one launch
    before = np.concatenate([[0], np.cumsum(need)])
    launches, first, count = [], 0, len(need)
    while first < count:
        # largest last with gap + sum(need[first:last]) <= max_rows
        last = int(np.searchsorted(
            before, before[first] + max_rows - gap, side='right')) - 1
        last = min(max(last, first + 1), count)
        launches.append(list(range(first, last)))
        first = last
    return launches


###############################################################################
# Execution
###############################################################################


def _prepare(audios, sample_rate):
    """A PackedAudio for a list of utterances"""
    if isinstance(audios, PackedAudio):
        if sample_rate != emphases.SAMPLE_RATE:
            raise ValueError('PackedAudio must already be at 16 kHz')
        return audios
    if sample_rate != emphases.SAMPLE_RATE:
        audios = [emphases.resample(audio, sample_rate) for audio in audios]
    cuda = [audio for audio in audios if audio.device.type == 'cuda']
    if cuda:
        audios = [audio.cpu() for audio in audios]
    return pack_audio(audios, pin=len(audios) > 1)


def run_on_device(
    model, alignments, audios, sample_rate, batch_size, device, to_cpu=True,
    output='scores'
):
    """Run a list of utterances on one device; returns a list of (1, W_i)
    score tensors (CPU when to_cpu, else on `device`).  output='logits'
    returns the network output before postprocessing (the evaluation caller,
    emphases/evaluate/core.py:73-94)."""
    if output not in ('scores', 'logits'):
        raise ValueError(f'output {output} is not defined')
    if model.location == 'input':
        return _run_via_model(
            model, alignments, audios, sample_rate, batch_size, device, to_cpu,
            output)
    packed = _prepare(audios, sample_rate)
    eng = emphases.get_engine(device)
    weights = model.packed_weights()
    method = emphases.DOWNSAMPLE_METHOD
    if method not in _lib.POOL:
        raise ValueError(f'Interpolation method {method} is not defined')
    if emphases.LOSS == 'bce':
        head_mode = _lib.HEAD_SIGMOID
    elif emphases.LOSS == 'mse':
        head_mode = _lib.HEAD_CLAMP
    else:
        head_mode = _lib.HEAD_LOGITS
    precision = emphases.precision_code()

    frames = (packed.lengths + 2 * engine.PADDING) // engine.HOPSIZE
    launches = bucket_launches(frames, emphases.MAX_ROWS_PER_LAUNCH)
    streams = [torch.cuda.current_stream(device)]
    if len(launches) > 1:
        streams = [torch.cuda.Stream(device) for _ in range(2)]
        for stream in streams:
            stream.wait_stream(torch.cuda.current_stream(device))
    pending = []
    for number, members in enumerate(launches):
        first, last = members[0], members[-1]
        base = int(packed.offsets[first])
        end = int(packed.offsets[last] + engine.align_samples(packed.lengths[last]))
        end = min(end, packed.buffer.numel())
        stream = streams[number % len(streams)]
        # one grow-only workspace per stream: launches on a stream run in
        # order, so its buffers can be reused without further synchronisation
        ws = eng.workspace(number % len(streams))
        if packed.ready is not None:
            packed.ready(last)
        with torch.cuda.stream(stream):
            # the audio copy goes out first: the host work below (alignment
            # conversion, planning) then overlaps it, and the copy engine never
            # waits for the host
            device_audio = ws.get('audio', (end - base,), packed.buffer.dtype)
            device_audio.copy_(packed.buffer[base:end], non_blocking=True)
        plan = engine.make_plan(
            [(as_times(alignments[i]), int(packed.lengths[i])) for i in members],
            batch_size,
            validate_method=method)
        with torch.cuda.stream(stream):
            result = eng.forward_packed(
                device_audio, plan, weights, method=method,
                location=model.location, precision=precision,
                head_mode=head_mode, normalize=emphases.NORMALIZE,
                views=eng.upload_plan(plan, slot=number, ws=ws), ws=ws)
            scores = result[output]
            if not to_cpu:
                scores = scores.clone()       # the workspace is reused
            if to_cpu:
                host = eng.pinned(
                    ('scores', number), scores.numel(), scores.dtype)
                host.copy_(scores, non_blocking=True)
                scores = host
        pending.append((members, plan, scores))
    for stream in streams:
        torch.cuda.current_stream(device).wait_stream(stream)
    if to_cpu:
        torch.cuda.current_stream(device).synchronize()

    outputs = [None] * len(alignments)
    for members, plan, scores in pending:
        # word rows without separators are all words of all utterances in order
        keep = torch.from_numpy(np.nonzero(plan.word_seq >= 0)[0])
        # (indexing copies, so the pinned staging buffer can be reused)
        flat = scores[keep.to(scores.device)] if len(keep) else scores[:0].clone()
        per_utterance = np.bincount(
            plan.utterance, weights=plan.n_words, minlength=len(members)
        ).astype(np.int64)
        pieces = torch.split(flat, per_utterance.tolist())
        for index, piece in zip(members, pieces):
            outputs[index] = piece[None]
    return outputs


def _run_via_model(
    model, alignments, audios, sample_rate, batch_size, device, to_cpu,
    output='scores'
):
    """Chunk-at-a-time path through Model.forward (transformer variant and
    the 'input' location), structured like the reference loop
    (emphases/core.py:245-265)"""
    outputs = []
    for alignment, audio in zip(alignments, audios):
        scores = []
        for features, bounds in emphases.preprocess(
            alignment, audio, sample_rate, batch_size, device.index
        ):
            logits = emphases.infer_with_model(model, features, bounds)[0]
            scores.append(
                emphases.postprocess(logits) if output == 'scores' else logits)
        result = torch.cat(scores, 1) if scores else torch.zeros(
            (1, 0), device=device)
        outputs.append(result.cpu() if to_cpu else result)
    return outputs


def run_sharded(alignments, audios, sample_rate, checkpoint, batch_size, gpus):
    """One worker thread per GPU over an LPT-balanced split of the corpus.
    Cost model: frames (conv) -- the transformer variant uses frames^2."""
    lengths = np.array([
        int(audios.lengths[i]) if isinstance(audios, PackedAudio)
        else int(audios[i].shape[-1]) for i in range(len(alignments))])
    rate = emphases.SAMPLE_RATE / float(sample_rate)
    frames = lengths * rate / engine.HOPSIZE
    costs = frames ** 2 if emphases.ARCHITECTURE == 'transformer' else frames
    shards = lpt_assign(costs.tolist(), len(gpus))
    outputs = [None] * len(alignments)
    errors = []

    def worker(gpu, shard):
        try:
            device = torch.device('cuda', gpu)
            with torch.cuda.device(device):
                # one model per device (load_model caches a single entry)
                model = _model_for(checkpoint, device)
                results = run_on_device(
                    model,
                    [alignments[i] for i in shard],
                    [audios[i] for i in shard],
                    sample_rate, batch_size, device, True)
            for index, result in zip(shard, results):
                outputs[index] = result
        except Exception as error:   # surfaced after join
            errors.append(error)

    threads = [
        threading.Thread(target=worker, args=(gpu, shard))
        for gpu, shard in zip(gpus, shards) if shard]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    if errors:
        raise errors[0]
    return outputs


_models = {}
_models_lock = threading.Lock()


def _model_for(checkpoint, device):
    with _models_lock:
        key = (str(checkpoint), device)
        if key not in _models:
            state = torch.load(
                checkpoint, map_location='cpu', weights_only=False)
            model = emphases.Model()
            model.load_state_dict(state['model'] if 'model' in state else state)
            _models[key] = model.to(device).eval()
        return _models[key]
