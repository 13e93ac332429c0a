"""Context-window scheduler: drop-in for the reference's ``pipelines/context.py`` (same names, same
results, bit-exact integers).

Reference: ``uniform`` pipelines/context.py:30-59, ``ordered_halving`` :22-27, ``get_context_scheduler``
:62-66, ``compute_num_context`` :7-10, ``compute_context_indices`` :13-19.  Host-side integer logic only.
"""
from typing import Callable, Iterator, List, Optional

import numpy as np

__all__ = ["ordered_halving", "uniform", "get_context_scheduler", "compute_num_context",
           "compute_context_indices", "get_total_steps", "window_table"]


def compute_num_context(init_video_length: int, context_size: int, context_overlap: int) -> int:
    stride = context_size - context_overlap
    return (init_video_length - context_size) // stride + 1


def compute_context_indices(num_context: int, context_size: int, context_overlap: int):
    stride = context_size - context_overlap
    return [(i * stride, i * stride + context_size - 1) for i in range(num_context)]


def ordered_halving(val: int) -> float:
    """Radical inverse in base 2 of a 64-bit integer (van der Corput), in [0, 1)."""
    v, out = int(val) & ((1 << 64) - 1), 0
    for _ in range(64):
        out = (out << 1) | (v & 1)
        v >>= 1
    return out / (1 << 64)


def uniform(step: int = ..., num_frames: int = ..., context_size: Optional[int] = None, context_stride: int = 3,
            context_overlap: int = 4, closed_loop: bool = True) -> Iterator[List[int]]:
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    n_dilations = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    frac = ordered_halving(step)
    shift = int(round(num_frames * frac))
    stop = num_frames + shift + (0 if closed_loop else -context_overlap)
    for level in range(n_dilations):
        dilation = 1 << level
        hop = context_size * dilation - context_overlap
        for start in range(int(frac * dilation) + shift, stop, hop):
            frames = []
            for e in range(start, start + context_size * dilation, dilation):
                # indices past the end are reflected back (the reference's `num_frames - 2 - e % num_frames`)
                frames.append(e if e < num_frames else num_frames - 2 - e % num_frames)
            yield frames


def get_context_scheduler(name: str) -> Callable:
    if name == "uniform":
        return uniform
    raise ValueError(f"Unknown context_overlap policy {name}")


def get_total_steps(scheduler, timesteps, num_steps=None, num_frames=..., context_size=None, context_stride=3,
                    context_overlap=4, closed_loop=True):
    return sum(len(list(scheduler(i, num_steps, num_frames, context_size, context_stride, context_overlap)))
               for i in range(len(timesteps)))


def window_table(video_length: int, context_frames: int, context_overlap: int, schedule: str = "uniform"):
    """The call the pipeline makes (reference pipelines/v_express_pipeline.py:486-500): windows for step 0 and
    the per-frame cover count.  The count reproduces the reference's NON-accumulating index-put: a frame that
    appears twice inside one (reflected) window is counted once for that window."""
    windows = list(get_context_scheduler(schedule)(step=0, num_frames=video_length, context_size=context_frames,
                                                   context_stride=1, context_overlap=context_overlap,
                                                   closed_loop=False))
    count = np.zeros(video_length, dtype=np.int64)
    for w in windows:
        count[np.unique(np.asarray(w, dtype=np.int64))] += 1
    return windows, count
