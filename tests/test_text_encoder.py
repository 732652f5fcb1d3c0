"""CLIP text encoder: lvd_b200.text_encoder.CLIPTextModel against transformers'
CLIPTextModel itself (the class the reference's pipeline loads; installed here as 5.5.0, the reference pins 4.36.2 --
same architecture) on identical random weights and token ids.  Gate as for the UNet:
err(engine fp16 vs HF fp32) <= max(3e-3, 1.5 x err(HF fp16 vs HF fp32))."""
import pytest
import torch

from conftest import rel_err


def _hf(cfg_kw, seed):
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(seed)
    m = CLIPTextModel(CLIPTextConfig(**cfg_kw)).eval()
    with torch.no_grad():  # default init is tiny (std 0.02): widen it so that attention / GELU actually matter
        for n, p in m.named_parameters():
            if p.dim() == 2 and "embedding" not in n:
                p.mul_(3.0)
            elif n.endswith("bias"):
                p.add_(0.1 * torch.randn_like(p))
    return m


def test_clip_param_spec_matches_transformers():
    from lvd_b200.text_encoder import clip_param_spec
    kw = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
              max_position_embeddings=77, hidden_act="gelu")
    sd = _hf(kw, 0).state_dict()
    spec = clip_param_spec(kw)
    ref = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith("position_ids")}
    assert spec == ref
    big = clip_param_spec(dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                               num_attention_heads=16, max_position_embeddings=77))
    assert sum(torch.Size(s).numel() for s in big.values()) == 340_387_840  # OpenCLIP ViT-H text tower, 23 layers


@pytest.mark.gpu
@pytest.mark.parametrize("kw,batch,seq", [
    (dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
          max_position_embeddings=77, hidden_act="gelu", eos_token_id=999), 3, 77),
    (dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
          max_position_embeddings=77, hidden_act="gelu", eos_token_id=999), 2, 9),
    (dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23, num_attention_heads=16,
          max_position_embeddings=77, hidden_act="gelu", eos_token_id=49407), 2, 77)])
def test_text_encoder_matches_transformers(cuda, kw, batch, seq):
    from lvd_b200.text_encoder import CLIPTextModel
    torch.backends.cuda.matmul.allow_tf32 = False
    hf = _hf(kw, 1)
    sd = {k: v.half() for k, v in hf.state_dict().items()}
    enc = CLIPTextModel(**{k: v for k, v in kw.items()}).to(cuda)
    enc.load_state_dict(sd)
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(1, kw["vocab_size"] - 1, (batch, seq), generator=g)
    for b in range(batch):  # EOS somewhere inside, padding (EOS id, as the OpenCLIP tokenizer pads) behind it
        e = 3 + 2 * b if seq > 9 else seq - 1 - b
        ids[b, e:] = kw["eos_token_id"]
    out = enc(ids.to(cuda))
    hf32 = hf.float().to(cuda)
    hf32.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
    with torch.no_grad():
        r32 = hf32(ids.to(cuda))
        r16 = hf32.half()(ids.to(cuda))
    e_h, e16 = rel_err(out[0], r32[0]), rel_err(r16[0], r32[0])
    e_p = rel_err(out.pooler_output, r32.pooler_output)
    print(f"\n[clip text {kw['num_hidden_layers']}x{kw['hidden_size']} B={batch} S={seq}] hidden engine-vs-fp32 {e_h:.3e}"
          f"   HF-fp16-vs-fp32 {e16:.3e}   pooled {e_p:.3e}")
    assert out[0].shape == (batch, seq, kw["hidden_size"]) and torch.isfinite(out[0].float()).all()
    assert e_h <= max(3e-3, 1.5 * e16) and e_p <= max(3e-3, 1.5 * e16)


@pytest.mark.gpu
def test_pipeline_encode_prompt_with_native_text_encoder(cuda):
    """_encode_prompt (controllable_pipeline_text_to_video_synth.py:197-372) with the native encoder and a stand-in
    tokenizer: cat([negative, positive]) embeddings of the right shape, equal to encoding the ids directly."""
    from lvd_b200.models.controllable_pipeline_text_to_video_synth import TextToVideoSDPipeline
    from lvd_b200.text_encoder import CLIPTextModel
    kw = dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
              max_position_embeddings=77, hidden_act="gelu", eos_token_id=999)
    enc = CLIPTextModel(**kw).to(cuda)
    enc.load_state_dict({k: v.half() for k, v in _hf(kw, 3).state_dict().items()})

    class Tok:
        model_max_length = 77

        def __call__(self, texts, padding=None, max_length=77, truncation=True, return_tensors="pt"):
            ids = torch.full((len(texts), 77), 999, dtype=torch.int64)
            for i, t in enumerate(texts):
                w = [1 + (hash(x) % 900) for x in t.split()][:75]
                ids[i, 0] = 998
                ids[i, 1:1 + len(w)] = torch.tensor(w, dtype=torch.int64)

            class R:
                input_ids = ids
            return R()

    class U:
        device = cuda

        class config:
            sample_size = 16
            in_channels = 4
            cross_attention_dim = 128

        def to(self, d):
            return self

    pipe = TextToVideoSDPipeline(unet=U(), text_encoder=enc, tokenizer=Tok())
    emb = pipe._encode_prompt("a bear walks left", cuda, 1, True, negative_prompt=None)
    assert emb.shape == (2, 77, 128) and emb.dtype == torch.float16
    direct = enc(Tok()(["", "a bear walks left"]).input_ids.to(cuda))[0]
    assert torch.equal(emb, direct)
