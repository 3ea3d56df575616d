"""Tiny, seeded, randomly initialised checkpoint directories laid out like the pipelines the reference loads
(code/models.py:478 AudioLDMPipeline, :556-564 AudioLDM2Pipeline, :402-422 TANGO snapshot) — test infrastructure for
the text-conditioning stage and the checkpoint loaders (no network: nothing here is a trained model).

    build_audioldm(dir)   tokenizer/ text_encoder/ (ClapTextModelWithProjection) unet/ scheduler/
    build_audioldm2(dir)  tokenizer/ text_encoder/ (ClapModel) tokenizer_2/ text_encoder_2/ (T5EncoderModel)
                          projection_model/ language_model/ (GPT2Model) unet/ scheduler/
    build_tango(dir)      main_config.json tokenizer/ pytorch_model_main.bin (text_encoder.* + unet.* keys)
"""
import json
import os

import torch

WORDS = ["a", "dog", "cat", "barking", "meowing", "recording", "of", "the", "piano", "music", "loud", "rain"]


def roberta_tokenizer(max_len=16):
    from tokenizers.pre_tokenizers import ByteLevel
    from transformers import RobertaTokenizer
    alpha = sorted(ByteLevel.alphabet())
    vocab = {t: i for i, t in enumerate(["<s>", "<pad>", "</s>", "<unk>"] + alpha + ["<mask>"])}
    return RobertaTokenizer(vocab=vocab, merges=[], model_max_length=max_len)


def t5_tokenizer(max_len=32):
    from transformers import T5Tokenizer
    pieces = [("<pad>", 0.0), ("</s>", 0.0), ("<unk>", 0.0), ("▁", -2.0)]
    pieces += [(c, -3.0) for c in "abcdefghijklmnopqrstuvwxyz"]
    pieces += [("▁" + w, -2.5) for w in WORDS]
    return T5Tokenizer(vocab=pieces, extra_ids=0, model_max_length=max_len)


def _clap_text_cfg(vocab_size):
    from transformers import ClapTextConfig
    return ClapTextConfig(vocab_size=vocab_size, hidden_size=32, num_hidden_layers=1, num_attention_heads=2,
                          intermediate_size=64, max_position_embeddings=20, projection_dim=24)


def _save_unet(dirname, preset, json_cfg, seed=0):
    """unet/config.json ([UPSTREAM] diffusers field names) + diffusion_pytorch_model.safetensors under diffusers
    state-dict names with the oracle's seeded synthetic weights."""
    from safetensors.torch import save_file
    from oracle import unet_torch as U
    from audioeditingcode_b200 import unet_config as C
    cfg = C.preset(preset)
    w = U.synthetic_weights(cfg, seed=seed)
    os.makedirs(os.path.join(dirname, "unet"), exist_ok=True)
    json.dump(json_cfg, open(os.path.join(dirname, "unet", "config.json"), "w"))
    save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(dirname, "unet", "diffusion_pytorch_model.safetensors"))
    return cfg, w


def _save_scheduler(dirname, beta_start, beta_end, pred):
    os.makedirs(os.path.join(dirname, "scheduler"), exist_ok=True)
    json.dump(dict(_class_name="DDIMScheduler", beta_start=beta_start, beta_end=beta_end, beta_schedule="scaled_linear",
                   prediction_type=pred, steps_offset=1, set_alpha_to_one=False, clip_sample=False,
                   timestep_spacing="leading", num_train_timesteps=1000),
              open(os.path.join(dirname, "scheduler", "scheduler_config.json"), "w"))


def build_audioldm(dirname, seed=0):
    from transformers import ClapTextModelWithProjection
    torch.manual_seed(seed)
    tok = roberta_tokenizer()
    tok.save_pretrained(os.path.join(dirname, "tokenizer"))
    ClapTextModelWithProjection(_clap_text_cfg(len(tok))).save_pretrained(os.path.join(dirname, "text_encoder"))
    cfg, w = _save_unet(dirname, "tiny-audioldm", dict(
        in_channels=8, out_channels=8, block_out_channels=[64, 128], layers_per_block=1,
        down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"], attention_head_dim=[2, 4],
        cross_attention_dim=[64, 128], class_embed_type="simple_projection", projection_class_embeddings_input_dim=512,
        class_embeddings_concat=True, norm_eps=1e-5, norm_num_groups=32), seed)
    _save_scheduler(dirname, 0.0015, 0.0195, "epsilon")
    return cfg, w


def build_audioldm2(dirname, seed=0):
    from safetensors.torch import save_file
    from transformers import ClapAudioConfig, ClapConfig, ClapModel, GPT2Config, GPT2Model, T5Config, T5EncoderModel
    torch.manual_seed(seed)
    tok = roberta_tokenizer()
    tok.save_pretrained(os.path.join(dirname, "tokenizer"))
    tok2 = t5_tokenizer()
    tok2.save_pretrained(os.path.join(dirname, "tokenizer_2"))
    ac = ClapAudioConfig(spec_size=64, patch_size=4, patch_stride=[4, 4], num_mel_bins=16, hidden_size=16, depths=[1, 1],
                         num_attention_heads=[1, 2], window_size=4, num_classes=4, patch_embeds_hidden_size=16)
    ClapModel(ClapConfig(text_config=_clap_text_cfg(len(tok)).to_dict(), audio_config=ac.to_dict(), projection_dim=24)
              ).save_pretrained(os.path.join(dirname, "text_encoder"))
    T5EncoderModel(T5Config(vocab_size=len(tok2), d_model=160, d_kv=16, d_ff=64, num_layers=1, num_heads=2,
                            feed_forward_proj="gated-gelu")).save_pretrained(os.path.join(dirname, "text_encoder_2"))
    gcfg = GPT2Config(vocab_size=8, n_positions=64, n_embd=96, n_layer=1, n_head=2)
    gcfg.max_new_tokens = 8
    GPT2Model(gcfg).save_pretrained(os.path.join(dirname, "language_model"))
    pm = os.path.join(dirname, "projection_model")
    os.makedirs(pm, exist_ok=True)
    json.dump(dict(_class_name="AudioLDM2ProjectionModel", text_encoder_dim=24, text_encoder_1_dim=160,
                   langauge_model_dim=96), open(os.path.join(pm, "config.json"), "w"))
    g = torch.Generator().manual_seed(seed + 1)
    save_file({"projection.weight": 0.2 * torch.randn(96, 24, generator=g), "projection.bias": 0.1 * torch.randn(96, generator=g),
               "projection_1.weight": 0.1 * torch.randn(96, 160, generator=g),
               "projection_1.bias": 0.1 * torch.randn(96, generator=g), "sos_embed": torch.randn(96, generator=g),
               "eos_embed": torch.randn(96, generator=g), "sos_embed_1": torch.randn(96, generator=g),
               "eos_embed_1": torch.randn(96, generator=g)}, os.path.join(pm, "diffusion_pytorch_model.safetensors"))
    cfg, w = _save_unet(dirname, "tiny-audioldm2", dict(
        in_channels=8, out_channels=8, block_out_channels=[64, 128], layers_per_block=1,
        down_block_types=["DownBlock2D", "CrossAttnDownBlock2D"], attention_head_dim=[2, 4],
        cross_attention_dim=[[None, 96, 160], [None, 96, 160]], norm_eps=1e-5, norm_num_groups=32), seed)
    _save_scheduler(dirname, 0.0015, 0.0195, "epsilon")
    return cfg, w


def build_tango(dirname, seed=0):
    from transformers import T5Config, T5EncoderModel
    from oracle import unet_torch as U
    from audioeditingcode_b200 import unet_config as C
    torch.manual_seed(seed)
    tok = t5_tokenizer()
    tok.save_pretrained(os.path.join(dirname, "tokenizer"))
    t5cfg = T5Config(vocab_size=len(tok), d_model=160, d_kv=16, d_ff=64, num_layers=1, num_heads=2,
                     feed_forward_proj="gated-gelu")
    enc_dir = os.path.join(dirname, "flan-t5-tiny")
    t5cfg.save_pretrained(enc_dir)
    enc = T5EncoderModel(t5cfg)
    json.dump(dict(text_encoder_name=enc_dir, scheduler_name="stabilityai/stable-diffusion-2-1",
                   unet_model_config_path="configs/diffusion_model_config.json"),
              open(os.path.join(dirname, "main_config.json"), "w"))
    cfg = C.preset("tiny-tango")
    w = U.synthetic_weights(cfg, seed=seed)
    sd = {"unet." + k: v for k, v in w.items()}
    sd.update({"text_encoder." + k: v for k, v in enc.state_dict().items()})
    torch.save(sd, os.path.join(dirname, "pytorch_model_main.bin"))
    return cfg, w, enc
