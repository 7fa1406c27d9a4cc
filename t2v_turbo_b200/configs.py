"""Model configurations of the VideoCrafter2 checkpoint that T2V-Turbo distils (plain dicts, accepted by
UNetModel / AutoencoderKL exactly like the reference's OmegaConf `params`)."""

# SURVEY.md Appendix A.1 = configs/inference_t2v_512_v2.0.yaml:24-50 + app.py:237-238
VC2_UNET = dict(
    in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], num_head_channels=64, transformer_depth=1, context_dim=1024, use_linear=True,
    use_checkpoint=False, temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
    use_relative_position=False, use_causal_attention=False, temporal_length=16, addition_attention=True,
    fps_cond=True, time_cond_proj_dim=256)

# configs/inference_t2v_512_v2.0.yaml:56-70
VC2_VAE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=512, in_channels=3, out_ch=3, ch=128,
                        ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)

