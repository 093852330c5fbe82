"""Host-side mirror of tinygpt::GPTModel / GPTEngine for the greedy decode path, over the C-ABI engine.

  DecodeEngine.forward(ids)      ↔ GPTModel::forward → model()(inputIds)            [ref: src/model/GPTModel.h:86]
  DecodeEngine.reset_cache()     ↔ GPTModel::resetCache                             [ref: src/model/GPTModel.h:91-94]
  DecodeEngine.gen_next_token()  ↔ GPTEngine::genNextToken (greedy sampler)         [ref: src/engine/GPTEngine.cpp:94-99]
  DecodeEngine.generate_sync()   ↔ GPTEngine::generateSync, batch 1, greedy         [ref: src/engine/GPTEngine.cpp:154-174]
  DecodeEngine.generate_async()  ↔ GPTEngine::generateAsync + AsyncTokenPipeline    [ref: src/engine/GPTEngine.cpp:17-35,180-232]

torch is used for device memory, pinned host buffers and the current stream only.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch

from . import models
from ._lib import B200Error, LayerWeights, ModelDesc, WeightTable, check, lib, require_device


def align_prompts(token_lists, context_size: int, pad_token: int):
    """GPTEngine::encodeTexts' alignment of a ragged batch [ref: src/engine/GPTEngine.cpp:101-141]: the common length is
    min(longest prompt, context size); longer prompts keep their LAST tokens, shorter ones are padded on the LEFT with
    `pad_token` (the tokenizer's pad id, else eos, else 0 — the caller resolves that).  → (ids int64 [B, S], mask bool
    [B, S] with False on the pads; the reference builds the mask and then ignores it, and so does the engine)."""
    if len(token_lists) == 0:
        raise B200Error("align_prompts: empty batch")
    S = min(max(len(t) for t in token_lists), int(context_size))
    ids = torch.full((len(token_lists), S), int(pad_token), dtype=torch.int64)
    mask = torch.zeros((len(token_lists), S), dtype=torch.bool)
    for i, t in enumerate(token_lists):
        t = list(t)[-S:] if S > 0 else []
        if t:
            ids[i, S - len(t):] = torch.tensor(t, dtype=torch.int64)
            mask[i, S - len(t):] = True
    return ids, mask


class DecodeEngine:
    """Whole-token engine for ONE sequence (batch 1) on the current CUDA device.

    `weights` is a dict with the reference's merged layout (models.synth_weights) already on the device; the engine
    borrows the tensors (keeps references) and owns only its KV cache + workspace."""

    def __init__(self, spec: models.ModelSpec, weights: Dict[str, torch.Tensor], rope_table: Optional[torch.Tensor] = None):
        require_device()
        self.spec = spec
        dev = weights["model.embed_tokens.weight"].device
        if dev.type != "cuda":
            raise B200Error("DecodeEngine: weights must be on a CUDA device (no CPU path)")
        self.device = dev
        self._w = weights  # keep alive
        if rope_table is None:
            # built on the device by the same fp32 powf/cosf/sinf kernels as the reference's op::ropeInit, so the table
            # is the one RoPE::cache() holds in a TinyGPT process (models.rope_table is the host/numpy restatement the
            # CPU oracle uses; libm and CUDA differ in the last bit of a few entries)
            from . import ops
            with torch.cuda.device(dev):
                rope_table = ops.rope_init(spec.head_dim, spec.max_ctx, spec.rope_theta, spec.rope_scaling, device=dev)
        self._rope = rope_table.to(dev).contiguous()
        self.local_vocab = spec.vocab
        desc, table = self._describe(spec, weights, rank=0, world=1, shard_attn=True)
        h = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib().b200_engine_create(C.byref(desc), C.byref(table), C.byref(h)), "b200_engine_create")
        self._h = h

    def _describe(self, spec, weights, rank: int, world: int, shard_attn: bool):
        """ModelDesc + WeightTable over `weights` (merged layout; for TP the tensors are this rank's shards)."""
        L = spec.layers
        arr = (LayerWeights * L)()
        for l in range(L):
            p = f"model.layers.{l}."
            g = lambda k: self._ptr(weights.get(p + k))
            arr[l] = LayerWeights(g("input_layernorm.weight"), g("self_attn.qkv_proj.weight"),
                                  g("self_attn.qkv_proj.bias"), g("self_attn.q_norm.weight"),
                                  g("self_attn.k_norm.weight"), g("self_attn.o_proj.weight"),
                                  g("post_attention_layernorm.weight"), g("mlp.gate_up_proj.weight"),
                                  g("mlp.down_proj.weight"))
        head = weights.get("lm_head.weight")
        if head is None:
            head = weights["model.embed_tokens.weight"]
        self._layers = arr
        table = WeightTable(self._ptr(weights["model.embed_tokens.weight"]), self._ptr(weights["model.norm.weight"]),
                            self._ptr(head), self._rope.data_ptr(), arr)
        desc = ModelDesc(spec.hidden, L, spec.q_heads, spec.kv_heads, spec.head_dim, spec.intermediate, spec.vocab,
                         spec.max_ctx, spec.rms_eps, int(spec.qkv_bias), int(spec.qk_norm), rank, world,
                         int(shard_attn))
        return desc, table

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
        if t is None:
            return None
        if t.dtype != torch.bfloat16 or not t.is_contiguous():
            raise B200Error("DecodeEngine: weights must be contiguous bf16 tensors")
        return t.data_ptr()

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib().b200_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------ state
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def rope_table(self) -> torch.Tensor:
        """fp32 [max_ctx, head_dim, 2] cos/sin table the engine rotates with."""
        return self._rope

    @property
    def position(self) -> int:
        return int(lib().b200_engine_position(self._h))

    @property
    def launches_per_token(self) -> int:
        return int(lib().b200_engine_launches_per_token(self._h))

    @property
    def options(self) -> dict:
        """Code paths this engine was built with (b200_engine_options)."""
        o = int(lib().b200_engine_options(self._h))
        return {"cuda_graph": bool(o & 1), "pdl": bool(o & 2), "gemm_prefill": bool(o & 8)}

    def bytes_per_token(self, ctx: int) -> int:
        return int(lib().b200_engine_bytes_per_token(self._h, ctx))

    def reset_cache(self) -> None:
        check(lib().b200_engine_reset(self._h, self._stream()), "b200_engine_reset")

    def seek(self, position: int) -> None:
        """Rewind to `position` inside the cached prefix (rows below it stay valid)."""
        check(lib().b200_engine_seek(self._h, int(position), self._stream()), "b200_engine_seek")

    # ---------------------------------------------------------------------------------------------- forward
    def forward(self, ids: torch.Tensor, all_positions: bool = False) -> torch.Tensor:
        """ids [1, S] int64 on the device → logits [1, S, V] (all_positions) or [1, 1, V] (last position) bf16.
        (Tensor-parallel engines return this rank's vocabulary shard, V / world columns.)"""
        if ids.dim() != 2 or not 1 <= ids.shape[0] <= 8:
            raise B200Error("DecodeEngine.forward: ids must be [B, S] with 1 ≤ B ≤ 8 (sequences at the same position)")
        if ids.dtype != torch.int64 or not ids.is_cuda:
            raise B200Error("DecodeEngine.forward: ids must be int64 on the device (reference: FuncNNLayer.h:205)")
        ids = ids.contiguous()
        B, S = ids.shape
        rows = S if all_positions else 1
        logits = torch.empty(B, rows, self.local_vocab, dtype=torch.bfloat16, device=self.device)
        check(lib().b200_engine_forward(self._h, ids.data_ptr(), B, S, logits.data_ptr(), int(all_positions),
                                        self._stream()), "b200_engine_forward")
        self._batch = B
        return logits

    def gen_next_token(self, ids: torch.Tensor) -> torch.Tensor:
        """genNextToken: forward → last position → greedy argmax (device tensor [B, 1] int64)."""
        B = ids.shape[0]
        check(lib().b200_engine_forward(self._h, ids.contiguous().data_ptr(), B, ids.shape[1], None, 0, self._stream()),
              "b200_engine_forward")
        self._batch = B
        out = torch.empty(B, 1, dtype=torch.int64, device=self.device)
        check(lib().b200_engine_last_token(self._h, out.data_ptr(), self._stream()), "b200_engine_last_token")
        return out

    def decode(self, n_steps: int) -> torch.Tensor:
        """n greedy steps on the device, each consuming the previous step's token → int64 [n] device tensor (after a
        batched forward with B > 1 sequences: [n, B])."""
        B = getattr(self, "_batch", 1)
        out = torch.empty((n_steps, B) if B > 1 else (n_steps,), dtype=torch.int64, device=self.device)
        check(lib().b200_engine_decode(self._h, n_steps, out.data_ptr(), self._stream()), "b200_engine_decode")
        return out

    def generate_sync_batch(self, prompts, max_new_tokens: int, pad_token: int = 0) -> torch.Tensor:
        """GPTEngine::generateSync for a batch of prompts with HOST buffers [ref: src/engine/GPTEngine.cpp:154-174]: one
        weight pass per step for the whole batch.  `prompts`: [B, S] ids, or a list of ragged id lists that is aligned
        the reference's way first (`align_prompts`: left padding with `pad_token`, truncation to the context; like the
        reference, the pad positions are attended — its mask is a TODO, GPTEngine.cpp:95).  → CPU int64 [B, new]."""
        if isinstance(prompts, (list, tuple)) and len(prompts) and not torch.is_tensor(prompts[0]) \
                and len({len(p) for p in prompts}) > 1:
            prompts, _ = align_prompts(prompts, self.spec.max_ctx, pad_token)
        prompts = torch.as_tensor(prompts, dtype=torch.int64)
        if prompts.dim() != 2 or prompts.shape[0] < 1:
            raise B200Error("generate_sync_batch: prompts must be [B, S]")
        ids = prompts.pin_memory().to(self.device, non_blocking=True)
        self.reset_cache()
        first = self.gen_next_token(ids)                       # [B, 1]
        rest = self.decode(max_new_tokens - 1)                 # [n-1, B] (or [n-1] for B = 1)
        rest = rest.view(max_new_tokens - 1, -1).t()
        toks = torch.cat([first, rest], dim=1).cpu()
        return toks

    # ------------------------------------------------------------------------------------------ generateSync
    def generate_sync(self, prompt_ids: List[int] | torch.Tensor, max_new_tokens: int,
                      pinned_in: Optional[torch.Tensor] = None, pinned_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """GPTEngine::generateSync for one sequence with HOST buffers: reset cache, H2D copy of the prompt ids,
        prefill, (max_new_tokens - 1) decode steps, ONE D2H copy of the generated ids (like the reference, no
        per-step host synchronisation).  Returns a CPU int64 tensor [max_new_tokens]."""
        if max_new_tokens < 1:
            raise B200Error("generate_sync: max_new_tokens must be >= 1")
        prompt = torch.as_tensor(prompt_ids, dtype=torch.int64).view(1, -1)
        if prompt.numel() == 0:
            raise B200Error("generate_sync: empty prompt")
        if pinned_in is None:
            pinned_in = torch.empty(prompt.shape, dtype=torch.int64).pin_memory()
        pinned_in.view(-1)[: prompt.numel()].copy_(prompt.view(-1))
        ids_dev = pinned_in.view(-1)[: prompt.numel()].to(self.device, non_blocking=True).view(1, -1)
        self.reset_cache()
        first = self.gen_next_token(ids_dev)
        rest = self.decode(max_new_tokens - 1)
        toks = torch.cat([first.view(-1), rest])
        if pinned_out is None:
            pinned_out = torch.empty(max_new_tokens, dtype=torch.int64).pin_memory()
        pinned_out[:max_new_tokens].copy_(toks, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return pinned_out[:max_new_tokens].clone()

    # ------------------------------------------------------------------------------------------------ sampler
    def set_sampler(self, temperature: float = 0.0, top_k: int = 0, top_p: float = 1.0, min_p: float = 0.0,
                    seed: int = 0) -> None:
        """SamplerConfig of the reference's generate loop (src/engine/Sampler.h:13-22).  Any knob set ⇒ every token is
        drawn by the device sampler with u = philox_uniform(seed, tokens generated so far); all off ⇒ greedy."""
        check(lib().b200_engine_set_sampler(self._h, float(temperature), int(top_k), float(top_p), float(min_p),
                                            int(seed) & 0xFFFFFFFFFFFFFFFF, self._stream()), "b200_engine_set_sampler")

    @staticmethod
    def philox_uniform(seed: int, n: int) -> float:
        """The uniform number the engine's sampler uses for its n-th generated token (n counts from 0): Philox4x32-10,
        counter {n lo, n hi, 0, 0}, key = seed; u = ((x0 >> 8) + 0.5)·2^-24.  Host mirror of sampling.cu."""
        M = 0xFFFFFFFF
        c = [n & M, (n >> 32) & M, 0, 0]
        k = [seed & M, (seed >> 32) & M]
        for _ in range(10):
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [((p1 >> 32) ^ c[1] ^ k[0]) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k[1]) & M, p0 & M]
            k = [(k[0] + 0x9E3779B9) & M, (k[1] + 0xBB67AE85) & M]
        return ((c[0] >> 8) + 0.5) / 16777216.0

    # ----------------------------------------------------------------------------------------- generateAsync
    def set_mailbox(self, capacity: int = 64) -> torch.Tensor:
        """Attach a token mailbox: a zeroed ring of `capacity` 8-byte words in pinned host memory that the argmax kernel
        posts {sequence tag, token} into (b200_engine_set_mailbox).  Returns the ring (int64 view of the words)."""
        ring = torch.zeros(capacity, dtype=torch.int64).pin_memory()
        check(lib().b200_engine_set_mailbox(self._h, ring.data_ptr(), capacity, self._stream()),
              "b200_engine_set_mailbox")
        self._ring, self._ring_np = ring, ring.numpy()  # numpy view: every element read is a fresh load
        return ring

    def clear_mailbox(self) -> None:
        check(lib().b200_engine_set_mailbox(self._h, None, 0, self._stream()), "b200_engine_set_mailbox")
        self._ring = self._ring_np = None

    def _fetch_token(self, n: int, timeout_s: float = 30.0) -> int:
        """Token number n (1-based count of tokens this engine generated) from the mailbox; spins on host memory."""
        cap = self._ring_np.shape[0]
        tag = n & 0xFFFFFFFF
        t0 = None
        while True:
            word = int(self._ring_np[(n - 1) % cap]) & 0xFFFFFFFFFFFFFFFF
            if (word >> 32) == tag:
                return word & 0xFFFFFFFF
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout_s:
                raise B200Error(f"generate_async: token {n} did not arrive within {timeout_s} s")

    def generate_async(self, prompt_ids, max_new_tokens: int, callback: Optional[Callable[[int], bool]] = None,
                       eos_ids: Iterable[int] = (), lookahead: int = 4) -> Tuple[List[int], str]:
        """GPTEngine::generateAsync for one sequence, greedy: tokens are handed to `callback(token_id) -> keep_going`
        as they are produced while the engine already runs up to `lookahead` steps ahead; no stream synchronisation or
        memcpy per token (the reference blocks on Tensor::item() for each one).  Stops at an EOS id (not passed to the
        callback, like the reference), when the callback returns False, or after max_new_tokens.
        Returns (token ids handed out, finish reason 'stop' | 'length').  Afterwards the engine is positioned so that
        feeding the next input token continues the sequence (steps that ran ahead of the stop are rewound)."""
        if max_new_tokens < 1:
            raise B200Error("generate_async: max_new_tokens must be >= 1")
        if getattr(self, "_ring_np", None) is None:
            self.set_mailbox(max(64, 4 * lookahead))
        if lookahead < 1 or lookahead >= self._ring_np.shape[0]:
            raise B200Error("generate_async: lookahead must be in [1, mailbox capacity)")
        eos = set(int(t) for t in eos_ids)
        prompt = torch.as_tensor(prompt_ids, dtype=torch.int64).view(1, -1)
        if prompt.numel() == 0:
            raise B200Error("generate_async: empty prompt")
        S = prompt.shape[1]
        self.reset_cache()
        base = self._generated()                               # tokens generated before this call
        self._enqueue_prefill(prompt)
        launched, out, reason = 1, [], "length"
        while len(out) < max_new_tokens:
            while launched < max_new_tokens and launched - len(out) < lookahead:
                self._enqueue_step()
                launched += 1
            tok = self._fetch_token(base + len(out) + 1)
            if tok in eos:
                reason = "stop"
                break
            out.append(tok)
            if callback is not None and not callback(tok):
                reason = "stop"
                break
        self._drain()
        # position of the next input token = S + (tokens kept) - 1 … the last kept token has not been fed yet unless a
        # look-ahead step consumed it; rewinding there makes both cases equal
        keep = max(len(out), 1)
        if S + keep - 1 < self.position:
            self.seek(S + keep - 1)
        return out, reason

    # the four touch points of generate_async with the device (overridden by the host-logic test's fake engine)
    def _generated(self) -> int:
        return int(lib().b200_engine_generated(self._h))

    def _enqueue_prefill(self, prompt: torch.Tensor) -> None:
        ids_dev = prompt.pin_memory().to(self.device, non_blocking=True)
        check(lib().b200_engine_forward(self._h, ids_dev.data_ptr(), 1, prompt.shape[1], None, 0, self._stream()),
              "b200_engine_forward")

    def _enqueue_step(self) -> None:
        check(lib().b200_engine_decode(self._h, 1, None, self._stream()), "b200_engine_decode")

    def _drain(self) -> None:
        torch.cuda.current_stream(self.device).synchronize()
