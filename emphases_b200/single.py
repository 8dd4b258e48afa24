"""One utterance through ONE native call (csrc/utterance.cu,
emph_infer_utterance): the latency path behind emphases.from_alignment_and_audio
(emphases/core.py:223-287) for the common case -- model-rate mono audio,
batch_size=None, the convolution architecture away from the 'input'
location.  Everything else (and any utterance the native planner declines)
goes through the batched scheduler, which produces the same scores."""
import ctypes
import threading

import numpy as np
import torch

import emphases_b200 as emphases
from . import _lib, engine
from .alignment import as_times


class _Runner:
    """Per (model weights, configuration) launch descriptor + grow-only workspace"""

    def __init__(self, model, device, weights, precision, method, head_mode):
        self.device = device
        eng = emphases.get_engine(device)
        frame = weights.frame
        frame_precision = engine.frame_precision(precision, frame)
        self.keep = [weights, eng]
        descriptor = _lib.UtteranceModel()

        def fill(stack_descriptor, stack, chosen):
            blob = stack.weights if chosen == _lib.PREC_FP32 \
                else stack.tensor_core_weights(chosen)
            acts = np.ascontiguousarray(stack.acts.astype(np.int32))
            self.keep += [blob, acts, stack.bias]
            stack_descriptor.weights = blob.data_ptr()
            stack_descriptor.bias = stack.bias.data_ptr()
            stack_descriptor.acts_host = acts.ctypes.data
            stack_descriptor.n_layers = stack.n_layers
            stack_descriptor.channels = stack.channels
            stack_descriptor.kernel_size = stack.kernel_size
            stack_descriptor.precision = chosen

        with torch.cuda.device(device), _lib.same_stream():
            fill(descriptor.frame, frame, frame_precision)
            descriptor.has_word_stack = int(model.location == 'intermediate')
            if descriptor.has_word_stack:
                fill(descriptor.word, weights.word,
                     engine.word_precision(precision, weights.word))
        descriptor.head_weight = weights.head_weight.data_ptr()
        descriptor.head_bias = float(weights.head_bias)
        descriptor.head_kernel = weights.head_kernel
        descriptor.head_mode = head_mode
        descriptor.mel_ptr = eng.mel_ptr.data_ptr()
        descriptor.mel_col = eng.mel_col.data_ptr()
        descriptor.mel_val = eng.mel_val.data_ptr()
        descriptor.n_mels = eng.n_mels
        descriptor.normalize = int(bool(emphases.NORMALIZE))
        descriptor.pool_method = _lib.POOL[method]
        self.descriptor = descriptor
        self.channels = frame.channels
        self.n_mels = eng.n_mels
        self.workspace = None
        self.logits = ctypes.c_void_p()
        self.scores = ctypes.c_void_p()
        self.lib = _lib.load()

    def run(self, times, audio, output):
        n_samples = audio.numel()
        n_words = len(times)
        need = self.lib.emph_infer_utterance_workspace(
            n_samples, n_words, self.channels, self.n_mels)
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = torch.empty(
                int(need * 1.25) + 4096, dtype=torch.uint8, device=self.device)
        workspace = self.workspace
        stream = torch._C._cuda_getCurrentRawStream(self.device.index)
        status = self.lib.emph_infer_utterance(
            ctypes.byref(self.descriptor), times.ctypes.data, n_words,
            audio.data_ptr(), int(audio.dtype == torch.int16), n_samples,
            workspace.data_ptr(), workspace.numel(),
            ctypes.byref(self.logits), ctypes.byref(self.scores), stream)
        if status == _lib.ENOSYS:
            return None
        if status != 0:
            message = self.lib.emph_last_error().decode('utf-8', 'replace')
            raise _lib.EmphasesB200Error(f'emph_infer_utterance failed ({status}): {message}')
        pointer = self.scores.value if output == 'scores' else self.logits.value
        offset = pointer - workspace.data_ptr()
        # (the workspace is reused by the next call: hand out a copy)
        return workspace[offset:offset + 4 * n_words].view(torch.float32).clone()[None]


_runners = {}
_lock = threading.Lock()


def from_alignment_and_audio(model, alignment, audio, sample_rate, device, output='scores'):
    """Scores (1, W) on `device`, or None when this utterance / configuration
    is not one the native single-utterance call handles"""
    if (
        sample_rate != emphases.SAMPLE_RATE or
        model.architecture != 'convolution' or
        model.location not in ('intermediate', 'loss', 'inference') or
        emphases.DOWNSAMPLE_METHOD not in _lib.POOL or
        not torch.is_tensor(audio) or
        audio.dtype != torch.float32 or
        not (audio.dim() == 1 or (audio.dim() == 2 and audio.shape[0] == 1)) or
        audio.numel() == 0 or not audio.is_contiguous() or
        (audio.device.type == 'cuda' and audio.device != device)
    ):
        return None
    emphases.require_mel_features_only()
    times = as_times(alignment)
    if times.dtype != np.float64 or not times.flags.c_contiguous or len(times) == 0:
        return None
    weights = model.packed_weights()
    if hasattr(weights, 'input_layer') or weights.frame.channels != emphases.NUM_MELS:
        return None
    precision = emphases.precision_code()
    method = emphases.DOWNSAMPLE_METHOD
    head_mode = (
        _lib.HEAD_SIGMOID if emphases.LOSS == 'bce' else
        _lib.HEAD_CLAMP if emphases.LOSS == 'mse' else _lib.HEAD_LOGITS)
    key = (id(weights), device, precision, method, head_mode, model.location,
           bool(emphases.NORMALIZE))
    runner = _runners.get(key)
    if runner is None:
        with _lock:
            # one entry per live weight set (a reloaded model replaces its key)
            for stale in [k for k in _runners if k[1] == device and k[0] != id(weights)]:
                del _runners[stale]
            runner = _runners[key] = _Runner(
                model, device, weights, precision, method, head_mode)
    return runner.run(times, audio, output)
