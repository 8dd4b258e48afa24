"""Corpus scheduling: launch bucketing on one GPU, length-balanced sharding
across GPUs.

Utterances are independent (emphases/core.py:169-179 loops over files), so a
corpus shards by utterance with no collective: each GPU runs its shard
through the packed kernels and the host gathers the per-word scores.
"""
import ctypes
import os
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

    def launch_source(self, number, first, last):
        """Host (or device) samples of utterances first..last, one contiguous
        slice -- the source of launch `number`'s single copy"""
        if self.ready is not None:
            self.ready(int(last))
        base = int(self.offsets[first])
        end = int(self.offsets[last] + engine.align_samples(self.lengths[last]))
        return self.buffer[base:min(end, self.buffer.numel())]

    def view(self, first, stop):
        """Utterances [first, stop) as a PackedAudio over the same buffer"""
        if stop <= first:
            return PackedAudio(self.buffer[:0], [], [])
        base = int(self.offsets[first])
        end = min(
            int(self.offsets[stop - 1] + engine.align_samples(self.lengths[stop - 1])),
            self.buffer.numel())
        part = PackedAudio(
            self.buffer[base:end], self.offsets[first:stop] - base,
            self.lengths[first:stop])
        if self.ready is not None:
            part.ready = lambda j, parent=self.ready: parent(first + int(j))
        return part

    @staticmethod
    def layout(lengths):
        lengths = np.asarray(lengths, dtype=np.int64)
        padded = engine.align_samples(lengths)
        offsets = np.concatenate([[0], np.cumsum(padded[:-1])]) \
            if len(lengths) else np.zeros(0, dtype=np.int64)
        total = int(padded.sum())
        return offsets.astype(np.int64), max(total, engine.AUDIO_ALIGN)


_staging_buffers = threading.local()


def staging(samples, dtype, pinned=True):
    """Grow-only host staging buffer of the calling thread (pinning costs
    ~15 ms per GB; repeated corpus calls reuse one buffer per dtype).  Valid
    until the same thread asks for the same dtype again."""
    cache = _staging_buffers.__dict__.setdefault('buffers', {})
    key = (dtype, bool(pinned))
    buffer = cache.get(key)
    if buffer is None or buffer.numel() < samples:
        buffer = torch.empty(
            max(int(samples * 1.25), 8), dtype=dtype,
            pin_memory=bool(pinned) and torch.cuda.is_available())
        cache[key] = buffer
    return buffer[:samples]


class StreamedPack(PackedAudio):
    """A list of per-utterance fp32 CPU tensors (what a caller of the
    reference holds, emphases/core.py:223-230) on its way into pinned
    staging: a background thread packs launch after launch on the native
    pool (emph_pack_audio_f32) while earlier launches upload and run.  A
    launch whose samples are all 16-bit PCM values (audio from load.audio) is
    staged as int16 -- same values, half the PCIe bytes."""

    def __init__(self, audios):
        rows = []
        for audio in audios:
            row = audio[0] if audio.dim() == 2 else audio
            if row.dtype != torch.float32 or not row.is_contiguous():
                row = row.to(torch.float32).contiguous()
            rows.append(row)
        self.rows = rows                                   # keeps the sources alive
        lengths = [int(row.shape[0]) for row in rows]
        offsets, total = PackedAudio.layout(lengths)
        super().__init__(staging(total, torch.float32), offsets, lengths)
        self.narrow_buffer = staging(total, torch.int16)
        self.pointers = np.array([row.data_ptr() for row in rows], dtype=np.uint64)
        self.narrowed = {}
        self.condition = threading.Condition()
        self.error = None
        self.worker = None

    def start(self, launches, threads=None):
        lib = _lib.load()
        threads = threads or min(16, os.cpu_count() or 1)

        def pack():
            try:
                for number, members in enumerate(launches):
                    first, count = members[0], len(members)
                    flag = ctypes.c_int32(0)
                    pointers = np.ascontiguousarray(self.pointers[first:first + count])
                    status = lib.emph_pack_audio_f32(
                        pointers.ctypes.data, self.lengths[first:].ctypes.data,
                        self.offsets[first:].ctypes.data, count,
                        ctypes.c_void_p(self.buffer.data_ptr()),
                        ctypes.c_void_p(self.narrow_buffer.data_ptr()),
                        ctypes.byref(flag), threads)
                    if status != 0:
                        raise _lib.EmphasesB200Error('emph_pack_audio_f32 failed')
                    with self.condition:
                        self.narrowed[number] = bool(flag.value)
                        self.condition.notify_all()
            except Exception as error:          # surfaced by launch_source
                with self.condition:
                    self.error = error
                    self.condition.notify_all()

        self.worker = threading.Thread(target=pack, daemon=True)
        self.worker.start()

    def launch_source(self, number, first, last):
        with self.condition:
            self.condition.wait_for(
                lambda: number in self.narrowed or self.error is not None)
            if self.error is not None:
                raise self.error
        base = int(self.offsets[first])
        end = int(self.offsets[last] + engine.align_samples(self.lengths[last]))
        source = self.narrow_buffer if self.narrowed[number] else self.buffer
        return source[base:min(end, source.numel())]

    def finish(self):
        if self.worker is not None:
            self.worker.join()
            self.worker = None


def pack_audio(audios, dtype=torch.float32, pin=True):
    """List of (C, T) tensors -> PackedAudio (channel 0, as mels.py:48)"""
    lengths = [int(audio.shape[-1]) for audio in audios]
    offsets, total = PackedAudio.layout(lengths)
    buffer = torch.zeros(
        total, dtype=dtype, pin_memory=pin and torch.cuda.is_available())
    for offset, audio in zip(offsets, audios):
        buffer[offset:offset + audio.shape[-1]] = audio[0]
    return PackedAudio(buffer, offsets, lengths)


class ResampledSource:
    """A packed corpus at another sample rate, converted on the fly: planning
    sees the model-rate lengths, every launch uploads the SOURCE samples of its
    utterances and the log-mel kernel resamples them tile by tile in shared
    memory (emph_logmel_resampled_*) -- the 16 kHz audio never exists in HBM
    (the reference resamples utterance by utterance first, core.py:613-619)."""

    def __init__(self, source, sample_rate):
        from . import resampling
        self.source = source              # PackedAudio / StreamedPack at `sample_rate`
        self.sample_rate = int(sample_rate)
        self.lengths = resampling.resampled_lengths(
            source.lengths, sample_rate, emphases.SAMPLE_RATE)

    def __len__(self):
        return len(self.lengths)

    def start(self, launches):
        if isinstance(self.source, StreamedPack):
            self.source.start(launches)

    def finish(self):
        if isinstance(self.source, StreamedPack):
            self.source.finish()

    def launch_source(self, number, first, last):
        return self.source.launch_source(number, first, last)

    def launch_resampling(self, members, plan, device):
        """Per-sequence source offsets (relative to the launch's slice) and
        lengths for emph_logmel_resampled_*"""
        from . import resampling
        first = members[0]
        base = int(self.source.offsets[first])
        utterances = np.asarray(members, dtype=np.int64)[plan.utterance]
        source_off = torch.from_numpy(
            (self.source.offsets[utterances] - base).astype(np.int64))
        source_len = torch.from_numpy(self.source.lengths[utterances].astype(np.int32))
        return {
            'bank': resampling.device_bank(self.sample_rate, emphases.SAMPLE_RATE, device),
            'source_off': source_off.to(device, non_blocking=True),
            'source_len': source_len.to(device, non_blocking=True)}


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


def snake_assign(costs, workers: int) -> List[List[int]]:
    """Length-balanced split for LONG lists (tens of thousands of files): items
    sorted by cost, dealt to the workers in a snake (0 .. W-1, W-1 .. 0, ...).
    Within a fraction of a percent of lpt_assign's balance at these sizes, but
    vectorised: lpt_assign's heap costs ~2 us per item on every rank."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind='stable')
    position = np.arange(len(order))
    lap, slot = position // workers, position % workers
    worker = np.where(lap % 2 == 0, slot, workers - 1 - slot)
    return [np.sort(order[worker == w]).tolist() for w in range(workers)]


def contiguous_assign(costs, workers: int) -> List[List[int]]:
    """Cut the item list into `workers` contiguous ranges of (nearly) equal
    total cost; returns, per worker, the item indices (possibly empty)"""
    costs = np.asarray(costs, dtype=np.float64)
    if len(costs) == 0:
        return [[] for _ in range(workers)]
    total = np.concatenate([[0.], np.cumsum(costs)])
    targets = total[-1] * np.arange(1, workers) / workers
    upper = np.clip(np.searchsorted(total, targets, side='left'), 1, len(costs))
    nearest = np.where(
        np.abs(total[upper - 1] - targets) <= np.abs(total[upper] - targets), upper - 1, upper)
    cuts = [0] + [int(c) for c in nearest] + [len(costs)]
    cuts = np.maximum.accumulate(np.clip(cuts, 0, len(costs)))
    return [list(range(int(a), int(b))) for a, b in zip(cuts[:-1], cuts[1:])]


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


def _prepare(audios, sample_rate, device=None):
    """A PackedAudio for a list of utterances"""
    if isinstance(audios, ResampledSource):
        return audios
    if isinstance(audios, PackedAudio):
        if sample_rate != emphases.SAMPLE_RATE:
            return ResampledSource(audios, sample_rate)
        return audios
    cuda = [audio for audio in audios if audio.device.type == 'cuda']
    if sample_rate != emphases.SAMPLE_RATE and (cuda or len(audios) < 2):
        audios = [emphases.resample(audio, sample_rate) for audio in audios]
        sample_rate = emphases.SAMPLE_RATE
    if cuda:
        audios = [audio.cpu() for audio in audios]
    if sample_rate != emphases.SAMPLE_RATE:
        # resampled inside the log-mel kernel, launch by launch
        return ResampledSource(StreamedPack(audios), sample_rate)
    if len(audios) > 1:
        return StreamedPack(audios)
    return pack_audio(audios, pin=False)


def _flat_result(members, plan, scores):
    """(utterance indices, their words' values in order as one flat tensor,
    words per utterance) of one launch"""
    # word rows without separators are all words of all utterances in order
    keep = torch.from_numpy(np.nonzero(plan.word_seq >= 0)[0])
    # (indexing copies, so the pinned staging buffer can be reused)
    flat_scores = scores[keep.to(scores.device)] if len(keep) else scores[:0].clone()
    per_utterance = np.bincount(
        plan.utterance, weights=plan.n_words, minlength=len(members)
    ).astype(np.int64)
    return members, flat_scores, per_utterance


# bench.py's host-path probe: run_on_device with everything but the kernels
# (decode, planning, uploads, score download, result hand-off): the time the
# host side of the files path needs on its own
HOST_PATH_PROBE = False


class _LaunchSink:
    """Hands every launch's flat host result to `consume` on a background
    thread as soon as its device -> host copy has landed, while later launches
    are still being decoded, uploaded and run (from_files_to_files writes the
    .pt files of launch i under the decode of launch i + k)."""

    def __init__(self, consume):
        import queue
        import threading
        self.consume = consume
        self.queue = queue.Queue()
        self.failure = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            item = self.queue.get()
            if item is None:
                return
            if self.failure is not None:
                continue                      # drain
            members, plan, scores, event = item
            try:
                event.synchronize()
                self.consume(*_flat_result(members, plan, scores))
            except BaseException as error:    # re-raised on the caller's thread
                self.failure = error

    def put(self, members, plan, scores, stream):
        event = torch.cuda.Event()
        event.record(stream)
        self.queue.put((members, plan, scores, event))

    def close(self):
        self.queue.put(None)
        self.thread.join()
        if self.failure is not None:
            raise self.failure


def run_on_device(
    model, alignments, audios, sample_rate, batch_size, device, to_cpu=True,
    output='scores', flat=False, sink=None
):
    """Run a list of utterances on one device; returns a list of (1, W_i)
    score tensors (CPU when to_cpu, else on `device`).  output='logits'
    returns the network output before postprocessing (the evaluation caller,
    emphases/evaluate/core.py:73-94).  flat=True (host results only) skips the
    per-utterance split: a list of (utterance indices, flat fp32 host tensor
    of their words in order, words per utterance), one entry per launch --
    what the native .pt writer of from_files_to_files consumes.  `sink`
    (with flat=True): a callable taking such an entry, called on a background
    thread launch by launch while later launches run; nothing is returned."""
    if sink is not None and not (flat and to_cpu):
        raise ValueError('a sink takes flat host results')
    if output not in ('scores', 'logits'):
        raise ValueError(f'output {output} is not defined')
    emphases.require_mel_features_only()
    if model.location == 'input':
        return _run_via_model(
            model, alignments, audios, sample_rate, batch_size, device, to_cpu,
            output)
    packed = _prepare(audios, sample_rate, device)
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
    # the tensor-core operand blobs are packed lazily and cached: do it here, on
    # the current stream, so that the side streams below (which wait on it)
    # never read a blob another stream is still packing
    eng.prepare_weights(weights, precision)

    frames = (packed.lengths + 2 * engine.PADDING) // engine.HOPSIZE
    launches = bucket_launches(frames, emphases.MAX_ROWS_PER_LAUNCH)
    if isinstance(packed, (StreamedPack, ResampledSource)):
        packed.start(launches)
    streams = [torch.cuda.current_stream(device)]
    if len(launches) > 1:
        streams = [torch.cuda.Stream(device) for _ in range(2)]
        for stream in streams:
            stream.wait_stream(torch.cuda.current_stream(device))
    pending = []
    consumer = _LaunchSink(sink) if sink is not None else None
    try:
        for number, members in enumerate(launches):
            first, last = members[0], members[-1]
            stream = streams[number % len(streams)]
            # one grow-only workspace per stream: launches on a stream run in
            # order, so its buffers can be reused without further synchronisation
            ws = eng.workspace(number % len(streams))
            source = packed.launch_source(number, first, last)
            with torch.cuda.stream(stream):
                # the audio copy goes out first: the host work below (alignment
                # conversion, planning) then overlaps it, and the copy engine never
                # waits for the host
                device_audio = ws.get(
                    f'audio_{source.dtype}', (source.numel(),), source.dtype)
                device_audio.copy_(source, non_blocking=True)
            plan = engine.make_plan(
                [(as_times(alignments[i]), int(packed.lengths[i])) for i in members],
                batch_size,
                validate_method=method)
            with torch.cuda.stream(stream):
                resample = packed.launch_resampling(members, plan, device) \
                    if isinstance(packed, ResampledSource) else None
                if HOST_PATH_PROBE:
                    eng.upload_plan(plan, slot=number, ws=ws)
                    scores = ws.get('probe_scores', (plan.total_word_rows,), torch.float32)
                else:
                    result = eng.forward_packed(
                        device_audio, plan, weights, method=method,
                        location=model.location, precision=precision,
                        head_mode=head_mode, normalize=emphases.NORMALIZE,
                        views=eng.upload_plan(plan, slot=number, ws=ws), ws=ws,
                        resample=resample)
                    scores = result[output]
                if not to_cpu:
                    scores = scores.clone()       # the workspace is reused
                if to_cpu:
                    host = eng.pinned(
                        ('scores', number), scores.numel(), scores.dtype)
                    host.copy_(scores, non_blocking=True)
                    scores = host
            if consumer is not None:
                consumer.put(members, plan, scores, stream)
            else:
                pending.append((members, plan, scores))
    except BaseException:
        # an error (e.g. a word without frames) must not leave the background
        # packer writing into staging buffers the next call reuses
        if isinstance(packed, (StreamedPack, ResampledSource)):
            packed.finish()
        torch.cuda.synchronize(device)
        if consumer is not None:
            try:
                consumer.close()
            except BaseException:
                pass
        raise
    for stream in streams:
        torch.cuda.current_stream(device).wait_stream(stream)
    if isinstance(packed, (StreamedPack, ResampledSource)):
        packed.finish()
        # the staging buffers are reused by this thread's next call
        torch.cuda.current_stream(device).synchronize()
    elif to_cpu:
        torch.cuda.current_stream(device).synchronize()

    if consumer is not None:
        consumer.close()
        return []
    outputs = [None] * len(alignments)
    launches_flat = []
    for members, plan, scores in pending:
        members, flat_scores, per_utterance = _flat_result(members, plan, scores)
        if flat:
            launches_flat.append((members, flat_scores, per_utterance))
            continue
        pieces = torch.split(flat_scores, per_utterance.tolist())
        for index, piece in zip(members, pieces):
            outputs[index] = piece[None]
    return launches_flat if flat else outputs


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
    packed = isinstance(audios, PackedAudio) and emphases.DOWNSAMPLE_LOCATION != 'input' \
        and emphases.ARCHITECTURE == 'convolution'
    if packed:
        # a packed corpus is cut into CONTIGUOUS ranges of equal cost: every
        # shard is a zero-copy view of the pinned buffer (an LPT shard would
        # have to be gathered utterance by utterance)
        shards = contiguous_assign(costs, len(gpus))
    else:
        shards = lpt_assign(costs.tolist(), len(gpus))
    outputs = [None] * len(alignments)
    errors = []

    def worker(gpu, shard):
        try:
            device = torch.device('cuda', gpu)
            with torch.cuda.device(device):
                # the same per-device cache emphases.infer uses
                model = emphases.load_model(checkpoint, device)
                results = run_on_device(
                    model,
                    [alignments[i] for i in shard],
                    audios.view(shard[0], shard[-1] + 1) if packed
                    else [audios[i] for i in shard],
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
