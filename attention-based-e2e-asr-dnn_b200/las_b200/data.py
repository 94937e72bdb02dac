"""Device-side collate + SpecAugment for the LAS training loader (reference src/utils.py:95-128).

    collate = DeviceCollator(device, use_specaug=True)
    loader = DataLoader(dataset, batch_size=..., collate_fn=collate)       # instead of dataset.collate_fn
    for x, y, lx, ly in loader: ...                                        # x already on the device, (B, T, F) fp32

Same contract as datasetTrainDev.collate_fn: utterances sorted by length (descending, stable like the reference's
`sorted(..., reverse=True)`), MFCCs padded with 0, transcripts with 29, `(mfccs, transcripts, mfcc_lens, transcript_lens)`
returned, SpecAugment = torchaudio FrequencyMasking(6) then TimeMasking(200) over the whole padded batch with ONE interval per
axis, drawn with the same `torch.rand(1)` calls in the same order as torchaudio.functional.mask_along_axis.  What changes
is where the work happens: the ragged frames cross the bus once, unpadded, and one kernel pads + masks on the device
(use it with num_workers=0: it touches CUDA).
"""
from __future__ import annotations

import torch
from torch.nn.utils.rnn import pad_sequence

from . import _lib
from ._lib import check, stream_ptr


def draw_mask_interval(mask_param: int, axis_len: int, p: float = 1.0):
    """torchaudio.functional.mask_along_axis's interval: two torch.rand(1) draws -> [start, end)."""
    if p != 1.0:                                              # torchaudio _get_mask_param: no clipping to the axis when p == 1.0,
        mask_param = min(mask_param, int(axis_len * p))       # so for T < mask_param the start can be negative (kept as is)
    if mask_param < 1:
        return 0, 0
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (axis_len - value)
    start = int(min_value.long())
    end = int(min_value.long() + value.long())
    return start, end


class DeviceCollator:
    def __init__(self, device, use_specaug: bool = False, freq_mask_param: int = 6, time_mask_param: int = 200,
                 mfcc_padding: float = 0.0, trans_padding: int = 29):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('DeviceCollator needs a CUDA device: las_b200 has no CPU fallback')
        self.use_specaug = use_specaug
        self.freq_mask_param, self.time_mask_param = freq_mask_param, time_mask_param
        self.mfcc_padding, self.trans_padding = mfcc_padding, trans_padding

    def __call__(self, batch):
        mfccs = [u[0] for u in batch]
        transcripts = [u[1] for u in batch]
        idx = sorted(range(len(mfccs)), key=lambda i: len(mfccs[i]), reverse=True)     # src/utils.py:105
        mfccs = [mfccs[i] for i in idx]
        transcripts = [transcripts[i] for i in idx]
        mfcc_lens = [len(m) for m in mfccs]
        transcript_lens = [len(t) for t in transcripts]
        B, T, F = len(mfccs), max(mfcc_lens), mfccs[0].shape[-1]
        frames = torch.cat([m.to(torch.float32) for m in mfccs], dim=0).pin_memory().to(self.device, non_blocking=True)
        offs, acc = [], 0
        for n in mfcc_lens:
            offs.append(acc)
            acc += n
        offsets = torch.tensor(offs, dtype=torch.int64).to(self.device, non_blocking=True)
        lens_dev = torch.tensor(mfcc_lens, dtype=torch.int32).to(self.device, non_blocking=True)
        f_lo = f_hi = t_lo = t_hi = 0
        if self.use_specaug:                                                             # freq first, then time (:121-124)
            f_lo, f_hi = draw_mask_interval(self.freq_mask_param, F)
            t_lo, t_hi = draw_mask_interval(self.time_mask_param, T)
        out = torch.empty(B, T, F, dtype=torch.float32, device=self.device)
        check(_lib.load().las_collate_specaug_f32(frames.data_ptr(), offsets.data_ptr(), lens_dev.data_ptr(), B, T, F,
                                                  float(self.mfcc_padding), f_lo, f_hi, t_lo, t_hi, 0.0, out.data_ptr(), stream_ptr()),
              'collate_specaug')
        ys = pad_sequence(transcripts, batch_first=True, padding_value=self.trans_padding)
        return out, ys, torch.tensor(mfcc_lens), torch.tensor(transcript_lens)
