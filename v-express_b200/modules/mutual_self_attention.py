"""``ReferenceAttentionControl`` for the B200 UNets: same constructor / ``update`` / ``clear`` API as the
reference (modules/mutual_self_attention.py:18-55, 321-387), without monkey-patching ``forward``.

The reference installs a closure on every transformer block and keeps the reference features in
``block.bank``.  Here the read-mode arithmetic (attn1 -> attn1_5(bank) -> attn2(audio) -> ff, :176-267) is
part of the fused kernel schedule of ``UNetEngine``; this class only (1) records the two attention weights on
the model (the closure captured them at hook-install time, :82-83), (2) pairs reader and writer blocks in the
reference's order -- stable sort by descending width over depth-first module order (:346-351) -- and
(3) hands the banks over with the CFG zero half prepended (:357-363).
"""
from typing import List

import torch

from .unet_2d_condition import UNet2DConditionModel
from .unet_3d import TemporalBasicTransformerBlock, UNet3DConditionModel, attention_block_order


def torch_dfs(model: torch.nn.Module):
    out = [model]
    for child in model.children():
        out += torch_dfs(child)
    return out


class ReferenceAttentionControl:
    def __init__(self, unet, mode="write", do_classifier_free_guidance=False, attention_auto_machine_weight=float("inf"),
                 gn_auto_machine_weight=1.0, style_fidelity=1.0, reference_attn=True, reference_adain=False,
                 fusion_blocks="midup", batch_size=1, reference_attention_weight=1., audio_attention_weight=1.,
                 reference_drop_rate=0.):
        assert mode in ["read", "write"]
        assert fusion_blocks in ["midup", "full"]
        self.unet = unet
        self.mode = mode
        self.reference_attn = reference_attn
        self.reference_adain = reference_adain
        self.fusion_blocks = fusion_blocks
        self.reference_attention_weight = reference_attention_weight
        self.audio_attention_weight = audio_attention_weight
        self.reference_drop_rate = reference_drop_rate
        if reference_drop_rate != 0.:
            raise ValueError("reference_drop_rate is a training-time option; inference uses 0")
        if isinstance(unet, UNet3DConditionModel):
            if mode != "read":
                raise ValueError("the B200 denoising UNet only implements the read side of the reference control")
            if fusion_blocks != "full":
                raise ValueError("the V-Express pipeline pairs all 16 blocks (fusion_blocks='full')")
            unet.reference_attention_weight = float(reference_attention_weight)
            unet.audio_attention_weight = float(audio_attention_weight)
            for blk in self._reader_blocks():
                blk.bank = []
        elif isinstance(unet, UNet2DConditionModel):
            if mode != "write":
                raise ValueError("the B200 ReferenceNet only implements the write side of the reference control")
            if fusion_blocks != "full":
                raise ValueError("the V-Express pipeline pairs all 16 blocks (fusion_blocks='full')")
            unet.write_banks = True
            for blk in unet.writer_blocks():
                blk.bank = []

    # ------------------------------------------------------------------
    def _reader_blocks(self) -> List[TemporalBasicTransformerBlock]:
        mods = dict(self.unet.named_modules())
        return [mods[n] for n in attention_block_order(self.unet)]

    @staticmethod
    def _writer_banks(writer) -> List[List[torch.Tensor]]:
        """Banks of the writer side in pairing order.  ``writer`` is either a reference-style control whose
        ``unet`` holds blocks with ``.bank`` and ``.norm1.normalized_shape`` (a ReferenceNet run in write mode),
        or any object exposing ``banks`` already in pairing order."""
        if hasattr(writer, "banks"):
            return [list(b) if isinstance(b, (list, tuple)) else [b] for b in writer.banks]
        blocks = [m for m in torch_dfs(writer.unet) if hasattr(m, "bank") and hasattr(m, "norm1")]
        blocks = sorted(blocks, key=lambda x: -x.norm1.normalized_shape[0])
        return [list(b.bank) for b in blocks]

    def update(self, writer, do_classifier_free_guidance=True, do_unconditional_forward=False, dtype=torch.float16):
        if not self.reference_attn:
            return
        readers = self._reader_blocks()
        banks = self._writer_banks(writer)
        if len(banks) != len(readers):
            raise ValueError(f"writer exposes {len(banks)} banks, reader has {len(readers)} blocks")
        if getattr(self.unet, "_engine", None) is not None:
            # new bank tensors may reuse the address / version / shape of the previous ones: never trust the cache key
            self.unet._engine._bank_cache.clear()
        for r, bank in zip(readers, banks):
            if do_classifier_free_guidance:
                r.bank = [torch.cat([torch.zeros_like(v), v]).to(dtype) for v in bank]
            elif do_unconditional_forward:
                r.bank = [torch.zeros_like(v).to(dtype) for v in bank]
            else:
                r.bank = [v.clone().to(dtype) for v in bank]

    def clear(self):
        if self.reference_attn and isinstance(self.unet, UNet2DConditionModel):
            for w in self.unet.writer_blocks():
                w.bank.clear()
        if self.reference_attn and isinstance(self.unet, UNet3DConditionModel):
            for r in self._reader_blocks():
                r.bank.clear()
            if self.unet._engine is not None:
                self.unet._engine._bank_cache.clear()      # persistent K/V buffers (and graphs) are kept
