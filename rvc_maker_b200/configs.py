"""Hyper-parameter presets of the reference (main/configs/v{1,2}/*.json ``model`` sections + data.sample_rate).

``ctor_args(cfg)`` gives the 18 positional values in the order of ``cpt["config"]`` (main/inference/train.py:730), so
``Synthesizer(*ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"])`` mirrors convert.py:564.
"""
_RB = dict(resblock="1", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])
_COMMON = dict(inter_channels=192, hidden_channels=192, filter_channels=768, n_heads=2, n_layers=6, kernel_size=3, p_dropout=0,
               upsample_initial_channel=512, spk_embed_dim=109, gin_channels=256, **_RB)

CONFIGS = {
    "v2_40k": dict(spec_channels=1025, segment_size=32, upsample_rates=[10, 10, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                   sr=40000, text_enc_hidden_dim=768, **_COMMON),
    "v2_48k": dict(spec_channels=1025, segment_size=36, upsample_rates=[12, 10, 2, 2], upsample_kernel_sizes=[24, 20, 4, 4],
                   sr=48000, text_enc_hidden_dim=768, **_COMMON),
    "v1_32k": dict(spec_channels=513, segment_size=40, upsample_rates=[10, 4, 2, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4, 4],
                   sr=32000, text_enc_hidden_dim=256, **_COMMON),
}

CTOR_ORDER = ["spec_channels", "segment_size", "inter_channels", "hidden_channels", "filter_channels", "n_heads", "n_layers",
              "kernel_size", "p_dropout", "resblock", "resblock_kernel_sizes", "resblock_dilation_sizes", "upsample_rates",
              "upsample_initial_channel", "upsample_kernel_sizes", "spk_embed_dim", "gin_channels", "sr"]


def ctor_args(cfg):
    return [cfg[k] for k in CTOR_ORDER]
