"""Weight-loader fast path (tinygpt_b200/loader.py) on the CPU: checkpoints written in the layout the reference's loader
reads (models.save_checkpoint — the same files oracle/_ref's harness fed to the reference's own SafeTensors.cpp when
the fixtures in tests/golden were made) come back bit-exact in the engine's merged layout, per-rank slices equal
tp.shard_weights of the full tensors, sharded (index) checkpoints load, and the reference's strict-load errors
(missing key, shape, dtype: src/util/SafeTensors.cpp:170-212) are raised."""
import json
import struct

import pytest
import torch

from tinygpt_b200 import engine, loader, models, tp

SPECS = [models.TINY_QWEN2, models.TINY_LLAMA, models.TINY_QWEN3, models.TINY_MISTRAL]


def _same(a, b):
    assert a.keys() == b.keys(), sorted(set(a) ^ set(b))
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k].view(torch.int16), b[k].view(torch.int16)), k


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: s.name)
def test_roundtrip_single_gpu(tmp_path, spec):
    w = models.synth_weights(spec, seed=3)
    models.save_checkpoint(spec, w, str(tmp_path))
    got_spec, got, report = loader.load_checkpoint(tmp_path, device="cpu")
    _same(got, w)
    assert report["unexpected"] == [] and report["files"] == 1
    for f in ("model_type", "hidden", "layers", "q_heads", "kv_heads", "head_dim", "intermediate", "vocab", "rope_theta",
              "rms_eps", "tie", "qkv_bias", "qk_norm", "max_ctx", "rope_scaling"):
        assert getattr(got_spec, f) == getattr(spec, f), f
    # every byte of the file's tensors was copied exactly once (tied head: none)
    assert report["bytes"] == sum(t.numel() * 2 for t in w.values())


@pytest.mark.parametrize("spec,shard_attn", [(models.TINY_MISTRAL, True), (models.TINY_QWEN2, True),
                                             (models.TINY_QWEN2, False), (models.TINY_QWEN3, True),
                                             (models.TINY_LLAMA, True)], ids=lambda v: getattr(v, "name", str(v)))
def test_per_rank_slices_equal_shard_weights(tmp_path, spec, shard_attn):
    w = models.synth_weights(spec, seed=5)
    models.save_checkpoint(spec, w, str(tmp_path))
    for rank in range(2):
        _, got, report = loader.load_checkpoint(tmp_path, device="cpu", rank=rank, world=2, shard_attn=shard_attn)
        _same(got, tp.shard_weights(spec, w, rank, 2, shard_attn))
        if spec.tie:  # the tied head is a view of the embedding, not a copy
            assert got["lm_head.weight"].data_ptr() == got["model.embed_tokens.weight"][rank * spec.vocab // 2:].data_ptr()


def _split_checkpoint(d):
    """Rewrite model.safetensors as two shards + model.safetensors.index.json."""
    raw = (d / "model.safetensors").read_bytes()
    (hlen,) = struct.unpack("<Q", raw[:8])
    header = json.loads(raw[8:8 + hlen])
    base = 8 + hlen
    names = sorted(header)
    parts = [names[::2], names[1::2]]
    wm = {}
    for i, part in enumerate(parts):
        h, blobs, off = {}, [], 0
        for n in part:
            s, e = header[n]["data_offsets"]
            h[n] = {"dtype": header[n]["dtype"], "shape": header[n]["shape"], "data_offsets": [off, off + e - s]}
            blobs.append(raw[base + s:base + e])
            off += e - s
            wm[n] = f"model-{i + 1:05d}-of-00002.safetensors"
        hj = json.dumps({"__metadata__": {"format": "pt"}, **h}).encode()
        with open(d / f"model-{i + 1:05d}-of-00002.safetensors", "wb") as f:
            f.write(struct.pack("<Q", len(hj)) + hj + b"".join(blobs))
    (d / "model.safetensors").unlink()
    (d / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": wm}))


def test_sharded_index_checkpoint(tmp_path):
    spec = models.TINY_QWEN3
    w = models.synth_weights(spec, seed=7)
    models.save_checkpoint(spec, w, str(tmp_path))
    _split_checkpoint(tmp_path)
    _, got, report = loader.load_checkpoint(tmp_path, device="cpu")
    _same(got, w)
    assert report["files"] == 2


def _rewrite_header(path, fn):
    raw = path.read_bytes()
    (hlen,) = struct.unpack("<Q", raw[:8])
    header = json.loads(raw[8:8 + hlen])
    fn(header)
    hj = json.dumps(header).encode()
    # offsets are relative to the end of the header, so the payload can stay as it is
    path.write_bytes(struct.pack("<Q", len(hj)) + hj + raw[8 + hlen:])


def test_strict_errors_like_the_reference(tmp_path):
    spec = models.TINY_MISTRAL
    w = models.synth_weights(spec, seed=1)
    models.save_checkpoint(spec, w, str(tmp_path))
    f = tmp_path / "model.safetensors"
    good = f.read_bytes()

    _rewrite_header(f, lambda h: h.pop("model.layers.1.mlp.up_proj.weight"))
    with pytest.raises(loader.LoaderError, match="Missing key: model.layers.1.mlp.up_proj.weight"):
        loader.load_checkpoint(tmp_path, device="cpu")
    f.write_bytes(good)
    _rewrite_header(f, lambda h: h["model.norm.weight"].update(dtype="F16"))
    with pytest.raises(loader.LoaderError, match="dtype not equal for tensor: model.norm.weight"):
        loader.load_checkpoint(tmp_path, device="cpu")
    f.write_bytes(good)
    _rewrite_header(f, lambda h: h["model.norm.weight"].update(shape=[spec.hidden // 2, 2]))
    with pytest.raises(loader.LoaderError, match="shape not equal for tensor: model.norm.weight"):
        loader.load_checkpoint(tmp_path, device="cpu")
    f.write_bytes(good)
    # an extra tensor is reported, and only fatal when asked for
    _rewrite_header(f, lambda h: h.update({"rotary_emb.inv_freq": dict(h["model.norm.weight"])}))
    _, _, report = loader.load_checkpoint(tmp_path, device="cpu")
    assert report["unexpected"] == ["rotary_emb.inv_freq"]
    with pytest.raises(loader.LoaderError, match="Unexpected key"):
        loader.load_checkpoint(tmp_path, device="cpu", strict_unexpected=True)
    f.write_bytes(good)
    # config errors
    cfg = json.loads((tmp_path / "config.json").read_text())
    (tmp_path / "config.json").write_text(json.dumps({**cfg, "model_type": "gpt2"}))
    with pytest.raises(loader.LoaderError, match="Unsupported model_type"):
        loader.load_checkpoint(tmp_path, device="cpu")
    (tmp_path / "config.json").write_text(json.dumps({**cfg, "torch_dtype": "float32"}))
    with pytest.raises(loader.LoaderError, match="torch_dtype"):
        loader.load_checkpoint(tmp_path, device="cpu")
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    with pytest.raises(loader.LoaderError):
        loader.load_checkpoint(tmp_path, device="cpu", rank=0, world=3)      # I = 448 does not shard 3 ways
    f.unlink()
    with pytest.raises(loader.LoaderError, match="Load model failed"):
        loader.load_checkpoint(tmp_path, device="cpu")


def test_config_defaults_follow_the_reference(tmp_path):
    """head_dim is DERIVED for llama / qwen2 / mistral even if config.json carries one (ModelLlama.h:37), rope_theta
    defaults differ per family (ModelConfig.cpp:84,88,96), llama3 scaling is honoured only for llama."""
    base = {"hidden_size": 256, "num_hidden_layers": 1, "num_attention_heads": 4, "num_key_value_heads": 2,
            "intermediate_size": 64, "vocab_size": 32, "torch_dtype": "bfloat16", "head_dim": 128}
    (tmp_path / "config.json").write_text(json.dumps({**base, "model_type": "llama"}))
    s = loader.load_model_config(tmp_path)
    assert (s.head_dim, s.rope_theta, s.rms_eps, s.tie, s.qkv_bias, s.qk_norm) == (64, 1.0, 1e-5, False, False, False)
    (tmp_path / "config.json").write_text(json.dumps({**base, "model_type": "qwen3", "rope_scaling": {"rope_type": "llama3"}}))
    s = loader.load_model_config(tmp_path)
    assert (s.head_dim, s.rope_theta, s.qk_norm, s.rope_scaling) == (128, 10000.0, True, None)
    (tmp_path / "config.json").write_text(json.dumps({**base, "model_type": "qwen2", "max_position_embeddings": 77}))
    s = loader.load_model_config(tmp_path)
    assert (s.head_dim, s.qkv_bias, s.max_ctx) == (64, True, 77) and loader.load_model_config(tmp_path, max_ctx=9).max_ctx == 9


@pytest.mark.parametrize("staging_bytes", [256, 3000, 1 << 16])
def test_chunked_staging_path(tmp_path, staging_bytes):
    """The staging-buffer path CUDA destinations use (chunks of rows, strided column slices, rows larger than the
    buffer) — forced onto a CPU destination with tiny buffers so that every branch runs without a GPU."""
    spec = models.TINY_LLAMA
    w = models.synth_weights(spec, seed=11)
    models.save_checkpoint(spec, w, str(tmp_path))
    _, got, _ = loader.load_checkpoint(tmp_path, device="cpu", staging_bytes=staging_bytes)
    _same(got, w)
    _, r1, _ = loader.load_checkpoint(tmp_path, device="cpu", rank=1, world=2, staging_bytes=staging_bytes)
    _same(r1, tp.shard_weights(spec, w, 1, 2))


# --------------------------------------------------------------------------------- loader fast path, CUDA destination
@pytest.mark.gpu
def test_loader_cuda_staging_path(built_lib, tmp_path):
    """The pinned-staging H2D path (chunked, strided column slices included) delivers the same bytes as the CPU path,
    and the loaded tensors drive the engine to the same tokens as the in-memory synthetic checkpoint."""
    from tinygpt_b200 import loader, tp
    spec = models.TINY_QWEN2
    w = models.synth_weights(spec, seed=9)
    models.save_checkpoint(spec, w, str(tmp_path))
    got_spec, got, _ = loader.load_checkpoint(tmp_path, device="cuda")
    for k, v in w.items():
        assert torch.equal(got[k].cpu().view(torch.int16), v.view(torch.int16)), k
    _, r1, _ = loader.load_checkpoint(tmp_path, device="cuda", rank=1, world=2)
    want = tp.shard_weights(spec, w, 1, 2)
    for k, v in want.items():
        assert torch.equal(r1[k].cpu().view(torch.int16), v.view(torch.int16)), k
    prompt = [3, 1, 4, 1, 5, 9, 2, 6, 5]
    a = engine.DecodeEngine(got_spec.with_ctx(128), got)
    b = engine.DecodeEngine(spec.with_ctx(128), {k: v.to("cuda") for k, v in w.items()})
    assert a.generate_sync(prompt, 16).tolist() == b.generate_sync(prompt, 16).tolist()
    a.close()
    b.close()


