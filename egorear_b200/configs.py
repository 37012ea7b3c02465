"""The hot-path hyper-parameters of the reference's shipped configs, as plain dicts.

Values are the `model_cfg` entries of configs/ego4view_{syn,rw}_heatmap_mvfex-n1_jqa*.yaml (mvf_cfg) and
configs/ego4view_{syn,rw}_pose3d*.yaml (pose3d_cfg); the LightningCLI surface is unchanged, these are only used
when the package is driven without the reference's YAML files (bench.py, smoke(), tests).
"""
import copy

_TRANSFORMER = dict(cross_attn_cfg=dict(num_heads=4, batch_first=True), spatial_attn_cfg=dict(num_heads=4, batch_first=True),
                    ffn_cfg=dict(feedforward_dims=512, num_fcs=2, ffn_drop=0.0))

MVF_CFG = dict(input_dims=128, embed_dims=256, num_former_layers=1, joint_query_adaptation=True,
               mvf_transformer_cfg=_TRANSFORMER)

ENCODER_CFG = dict(resnet_cfg=dict(model_name="resnet18", out_stride=4, use_imagenet_pretrain=False),
                   neck_cfg=dict(in_channels=[64, 128, 256, 512], out_channels=128))

POSE3D_CFG = dict(num_joints=16, input_dims=128, embed_dims=128, mlp_dims=1024, mlp_dropout=0.0, num_mlp_layers=2,
                  num_former_layers=3, num_pred_mlp_layers=2, feat_down_stride=4, norm_mlp_pred=False, coor_norm_max=None,
                  coor_norm_min=None, conv_heatmap_dim_init=32, use_mlp_avgpool=False, use_mlp_heatmap=False,
                  camera_calib_file_dir_path=None, transformer_cfg=_TRANSFORMER)


def heatmap_mvfex_cfg(num_views=4, camera_model="ego4view_syn", **over):
    cfg = dict(num_views=num_views, image_size=[256, 256], num_heatmap=15, feat_down_stride=4, heatmap_threshold=0.5,
               encoder_cfg=copy.deepcopy(ENCODER_CFG), mvf_cfg=copy.deepcopy(MVF_CFG), camera_model=camera_model)
    cfg.update(over)
    return cfg


def pose3d_cfg(num_views=4, camera_model="ego4view_syn", use_pred_heatmap_init=True, **over):
    cfg = dict(num_views=num_views, image_size=[256, 256], use_pred_heatmap_init=use_pred_heatmap_init,
               camera_model=camera_model, **copy.deepcopy(POSE3D_CFG))
    cfg.update(over)
    return cfg
