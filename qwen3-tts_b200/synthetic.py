"""Seeded random weights with the reference's state_dict names and the expected shipped shapes.
There are no checkpoints offline: benchmarks, smoke tests and parity tests run on these."""
import math

import torch

from .config import CodecConfig, StackConfig, TTSConfig


def cfg_1p7b() -> TTSConfig:
    return TTSConfig(talker=StackConfig(2048, 28, 16, 8, 128, 6144, 3072, 1e-6, 1e6),
                     cp=StackConfig(1024, 5, 16, 8, 128, 3072, 2048, 1e-6, 1e6))


def cfg_0p6b() -> TTSConfig:
    return TTSConfig(talker=StackConfig(1024, 28, 16, 8, 128, 3072, 3072, 1e-6, 1e6),
                     cp=StackConfig(1024, 5, 16, 8, 128, 3072, 2048, 1e-6, 1e6))


def cfg_tiny() -> TTSConfig:
    return TTSConfig(talker=StackConfig(256, 3, 4, 2, 128, 512, 3072, 1e-6, 1e6),
                     cp=StackConfig(128, 2, 4, 2, 128, 256, 2048, 1e-6, 1e4), text_hidden_size=256,
                     text_vocab_size=1000, tts_bos_token_id=997, tts_eos_token_id=998, tts_pad_token_id=996)


def random_tts_weights(cfg: TTSConfig, device="cuda:0", seed=0, dtype=torch.bfloat16, std=0.02, with_text=False,
                       text_vocab=None):
    g = torch.Generator(device=device).manual_seed(seed)
    W = {}

    def rn(*shape, s=std):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s).to(dtype)

    def norm(n):
        return (1.0 + 0.1 * torch.randn(n, generator=g, device=device, dtype=torch.float32)).to(dtype)

    def stack(pfx, c):
        for i in range(c.num_layers):
            p = f"{pfx}.layers.{i}"
            W[f"{p}.self_attn.q_proj.weight"] = rn(c.num_heads * c.head_dim, c.hidden_size)
            W[f"{p}.self_attn.k_proj.weight"] = rn(c.num_kv_heads * c.head_dim, c.hidden_size)
            W[f"{p}.self_attn.v_proj.weight"] = rn(c.num_kv_heads * c.head_dim, c.hidden_size)
            W[f"{p}.self_attn.o_proj.weight"] = rn(c.hidden_size, c.num_heads * c.head_dim)
            W[f"{p}.self_attn.q_norm.weight"] = norm(c.head_dim)
            W[f"{p}.self_attn.k_norm.weight"] = norm(c.head_dim)
            W[f"{p}.mlp.gate_proj.weight"] = rn(c.intermediate_size, c.hidden_size)
            W[f"{p}.mlp.up_proj.weight"] = rn(c.intermediate_size, c.hidden_size)
            W[f"{p}.mlp.down_proj.weight"] = rn(c.hidden_size, c.intermediate_size)
            W[f"{p}.input_layernorm.weight"] = norm(c.hidden_size)
            W[f"{p}.post_attention_layernorm.weight"] = norm(c.hidden_size)
        W[f"{pfx}.norm.weight"] = norm(c.hidden_size)

    t, c = cfg.talker, cfg.cp
    stack("talker.model", t)
    W["talker.model.codec_embedding.weight"] = rn(t.vocab_size, t.hidden_size)
    W["talker.codec_head.weight"] = rn(t.vocab_size, t.hidden_size, s=0.05)
    stack("talker.code_predictor.model", c)
    for j in range(cfg.num_code_groups - 1):
        W[f"talker.code_predictor.model.codec_embedding.{j}.weight"] = rn(c.vocab_size, t.hidden_size)
        W[f"talker.code_predictor.lm_head.{j}.weight"] = rn(c.vocab_size, c.hidden_size, s=0.05)
    if c.hidden_size != t.hidden_size:
        W["talker.code_predictor.small_to_mtp_projection.weight"] = rn(c.hidden_size, t.hidden_size)
        W["talker.code_predictor.small_to_mtp_projection.bias"] = rn(c.hidden_size, s=0.01)
    if with_text:
        tv = text_vocab or cfg.text_vocab_size
        W["talker.model.text_embedding.weight"] = rn(tv, cfg.text_hidden_size)
        W["talker.text_projection.linear_fc1.weight"] = rn(cfg.text_hidden_size, cfg.text_hidden_size)
        W["talker.text_projection.linear_fc1.bias"] = rn(cfg.text_hidden_size, s=0.01)
        W["talker.text_projection.linear_fc2.weight"] = rn(t.hidden_size, cfg.text_hidden_size)
        W["talker.text_projection.linear_fc2.bias"] = rn(t.hidden_size, s=0.01)
    return W


def random_codec_weights(cfg: CodecConfig, device="cuda:0", seed=0, dtype=torch.float32):
    """Scales keep activations O(1) through the 1920x upsampling stack so the waveform is neither silent nor
    saturated (same recipe as oracle/codec.py:random_weights, restated — the product never imports oracle/)."""
    g = torch.Generator(device=device).manual_seed(seed)
    W = {}

    def rn(*shape, s=1.0):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s).to(dtype)

    def conv(name, cout, cin, k, groups=1, transposed=False):
        fan = cin * max(1, k // 2) if transposed else (cin // groups) * k
        shape = (cin, cout, k) if transposed else (cout, cin // groups, k)
        W[f"{name}.weight"] = rn(*shape, s=1.0 / math.sqrt(fan))
        W[f"{name}.bias"] = rn(cout, s=0.02)

    def lin(name, o, i, bias=False):
        W[f"{name}.weight"] = rn(o, i, s=1 / math.sqrt(i))
        if bias:
            W[f"{name}.bias"] = rn(o, s=0.02)

    half = cfg.codebook_dim // 2
    for pfx, n in (("quantizer.rvq_first", 1), ("quantizer.rvq_rest", cfg.num_quantizers - 1)):
        for i in range(n):
            W[f"{pfx}.vq.layers.{i}._codebook.embedding_sum"] = rn(cfg.codebook_size, half, s=0.5)
            W[f"{pfx}.vq.layers.{i}._codebook.cluster_usage"] = \
                (torch.rand(cfg.codebook_size, generator=g, device=device) + 0.5).to(dtype)
        W[f"{pfx}.input_proj.weight"] = rn(half, cfg.codebook_dim, 1, s=1 / math.sqrt(cfg.codebook_dim))
        W[f"{pfx}.output_proj.weight"] = rn(cfg.codebook_dim, half, 1, s=1 / math.sqrt(half))
    conv("pre_conv.conv", cfg.latent_dim, cfg.codebook_dim, 3)
    p, Hh = "pre_transformer", cfg.hidden_size
    lin(f"{p}.input_proj", Hh, cfg.latent_dim, True)
    lin(f"{p}.output_proj", cfg.latent_dim, Hh, True)
    one = lambda n, m, s: (m + s * torch.randn(n, generator=g, device=device)).to(dtype)  # noqa: E731
    for i in range(cfg.num_layers):
        lp = f"{p}.layers.{i}"
        lin(f"{lp}.self_attn.q_proj", cfg.num_heads * cfg.head_dim, Hh)
        lin(f"{lp}.self_attn.k_proj", cfg.num_kv_heads * cfg.head_dim, Hh)
        lin(f"{lp}.self_attn.v_proj", cfg.num_kv_heads * cfg.head_dim, Hh)
        lin(f"{lp}.self_attn.o_proj", Hh, cfg.num_heads * cfg.head_dim)
        lin(f"{lp}.mlp.gate_proj", cfg.intermediate_size, Hh)
        lin(f"{lp}.mlp.up_proj", cfg.intermediate_size, Hh)
        lin(f"{lp}.mlp.down_proj", Hh, cfg.intermediate_size)
        W[f"{lp}.input_layernorm.weight"] = one(Hh, 1.0, 0.1)
        W[f"{lp}.post_attention_layernorm.weight"] = one(Hh, 1.0, 0.1)
        W[f"{lp}.self_attn_layer_scale.scale"] = one(Hh, 0.3, 0.05)
        W[f"{lp}.mlp_layer_scale.scale"] = one(Hh, 0.3, 0.05)
    W[f"{p}.norm.weight"] = one(Hh, 1.0, 0.1)
    Cc = cfg.latent_dim
    for i, f in enumerate(cfg.upsampling_ratios):
        conv(f"upsample.{i}.0.conv", Cc, Cc, f, transposed=True)
        q = f"upsample.{i}.1"
        conv(f"{q}.dwconv.conv", Cc, Cc, 7, groups=Cc)
        W[f"{q}.norm.weight"] = one(Cc, 1.0, 0.1)
        W[f"{q}.norm.bias"] = rn(Cc, s=0.02)
        lin(f"{q}.pwconv1", 4 * Cc, Cc, True)
        lin(f"{q}.pwconv2", Cc, 4 * Cc, True)
        W[f"{q}.gamma"] = one(Cc, 0.3, 0.05)
    conv("decoder.0.conv", cfg.decoder_dim, Cc, 7)
    for i, r in enumerate(cfg.upsample_rates):
        cin, cout = cfg.decoder_dim // 2 ** i, cfg.decoder_dim // 2 ** (i + 1)
        bp = f"decoder.{i + 1}.block"
        W[f"{bp}.0.alpha"] = rn(cin, s=0.3)
        W[f"{bp}.0.beta"] = rn(cin, s=0.3)
        conv(f"{bp}.1.conv", cout, cin, 2 * r, transposed=True)
        for u in range(3):
            q = f"{bp}.{u + 2}"
            W[f"{q}.act1.alpha"] = rn(cout, s=0.3)
            W[f"{q}.act1.beta"] = rn(cout, s=0.3)
            conv(f"{q}.conv1.conv", cout, cout, 7)
            W[f"{q}.act2.alpha"] = rn(cout, s=0.3)
            W[f"{q}.act2.beta"] = rn(cout, s=0.3)
            conv(f"{q}.conv2.conv", cout, cout, 1)
            W[f"{q}.conv1.conv.weight"] *= 0.5
            W[f"{q}.conv2.conv.weight"] *= 0.5
    n = len(cfg.upsample_rates)
    cl = cfg.decoder_dim // 2 ** n
    W[f"decoder.{n + 1}.alpha"] = rn(cl, s=0.3)
    W[f"decoder.{n + 1}.beta"] = rn(cl, s=0.3)
    conv(f"decoder.{n + 2}.conv", 1, cl, 7)
    W[f"decoder.{n + 2}.conv.weight"] *= 0.3
    return W


def cfg_encoder_tiny():
    from .config import EncoderConfig
    return EncoderConfig(num_filters=8, hidden_size=64, num_layers=2, num_heads=4, head_dim=16, intermediate_size=96,
                         sliding_window=6, codebook_size=64, codebook_dim=32)


def random_encoder_weights(cfg, seed=0):
    """Seeded fp32 weights under MimiModel's state_dict names (encoder half), CPU tensors (same recipe as
    oracle/mimi_encoder.py:random_weights, restated — the product never imports oracle/)."""
    import math
    g = torch.Generator().manual_seed(seed)
    W = {}

    def rn(*shape, s=1.0):
        return torch.randn(*shape, generator=g) * s

    def conv(name, cin, cout, k):
        W[f"{name}.conv.weight"] = rn(cout, cin, k, s=1.0 / math.sqrt(cin * k))
        W[f"{name}.conv.bias"] = rn(cout, s=0.05)

    idx = 0
    conv(f"encoder.layers.{idx}", 1, cfg.num_filters, cfg.kernel_size)
    idx += 1
    dim = cfg.num_filters
    for r in cfg.ratios:
        conv(f"encoder.layers.{idx}.block.1", dim, dim // cfg.compress, cfg.residual_kernel_size)
        conv(f"encoder.layers.{idx}.block.3", dim // cfg.compress, dim, 1)
        idx += 2
        conv(f"encoder.layers.{idx}", dim, dim * 2, 2 * r)
        idx += 1
        dim *= 2
    idx += 1
    conv(f"encoder.layers.{idx}", dim, cfg.hidden_size, cfg.last_kernel_size)
    C, I = cfg.hidden_size, cfg.intermediate_size
    for l in range(cfg.num_layers):
        p = f"encoder_transformer.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            W[p + f"self_attn.{n}.weight"] = rn(C, C, s=1.0 / math.sqrt(C))
        W[p + "mlp.fc1.weight"] = rn(I, C, s=1.0 / math.sqrt(C))
        W[p + "mlp.fc2.weight"] = rn(C, I, s=1.0 / math.sqrt(I))
        for n in ("input_layernorm", "post_attention_layernorm"):
            W[p + n + ".weight"] = 1.0 + rn(C, s=0.1)
            W[p + n + ".bias"] = rn(C, s=0.05)
        W[p + "self_attn_layer_scale.scale"] = 0.3 + rn(C, s=0.05)
        W[p + "mlp_layer_scale.scale"] = 0.3 + rn(C, s=0.05)
    W["downsample.conv.weight"] = rn(C, C, 2 * cfg.downsample_stride, s=1.0 / math.sqrt(C * 4))
    D, K = cfg.codebook_dim, cfg.codebook_size
    nsem = cfg.num_semantic_quantizers
    for which, n in (("semantic", nsem), ("acoustic", cfg.valid_num_quantizers - nsem)):
        p = f"quantizer.{which}_residual_vector_quantizer."
        W[p + "input_proj.weight"] = rn(D, C, 1, s=1.0 / math.sqrt(C))
        for qi in range(n):
            usage = torch.rand(K, generator=g) * 3 + 0.5
            W[p + f"layers.{qi}.codebook.cluster_usage"] = usage
            W[p + f"layers.{qi}.codebook.embed_sum"] = rn(K, D, s=0.8 * (0.75 ** qi)) * usage[:, None]
    return W


def cfg_speaker_encoder_tiny():
    from .config import SpeakerEncoderConfig
    return SpeakerEncoderConfig(mel_dim=16, enc_dim=24, enc_channels=(32, 32, 32, 64), enc_kernel_sizes=(5, 3, 3, 1),
                                enc_dilations=(1, 2, 3, 1), enc_attention_channels=8, enc_res2net_scale=4, enc_se_channels=8)


def random_speaker_encoder_weights(cfg, seed=0):
    """Seeded fp32 weights under the reference's `speaker_encoder.` state_dict names (prefix stripped), CPU tensors."""
    import math
    g = torch.Generator().manual_seed(seed)
    W = {}

    def conv(p, cin, cout, k):
        W[p + ".weight"] = torch.randn(cout, cin, k, generator=g) / math.sqrt(cin * k)
        W[p + ".bias"] = torch.randn(cout, generator=g) * 0.05

    ch, ks, s = cfg.enc_channels, cfg.enc_kernel_sizes, cfg.enc_res2net_scale
    conv("blocks.0.conv", cfg.mel_dim, ch[0], ks[0])
    for i in range(1, len(ch) - 1):
        p = f"blocks.{i}"
        conv(p + ".tdnn1.conv", ch[i - 1], ch[i], 1)
        for j in range(s - 1):
            conv(f"{p}.res2net_block.blocks.{j}.conv", ch[i] // s, ch[i] // s, ks[i])
        conv(p + ".tdnn2.conv", ch[i], ch[i], 1)
        conv(p + ".se_block.conv1", ch[i], cfg.enc_se_channels, 1)
        conv(p + ".se_block.conv2", cfg.enc_se_channels, ch[i], 1)
    conv("mfa.conv", ch[-1], ch[-1], ks[-1])
    conv("asp.tdnn.conv", ch[-1] * 3, cfg.enc_attention_channels, 1)
    conv("asp.conv", cfg.enc_attention_channels, ch[-1], 1)
    conv("fc", ch[-1] * 2, cfg.enc_dim, 1)
    return W
