"""Model blocks of the ViTPose configs this path serves, as plain dicts.

They are the ``model = dict(...)`` blocks of the reference's config files, e.g.
configs/body/2d_kpt_sview_rgb_img/topdown_heatmap/coco/ViTPose_base_coco_256x192.py:52-84 and
configs/wholebody/.../coco-wholebody/ViTPose_huge_wholebody_256x192.py:29-59 — a reference config
file loaded with mmcv.Config gives the same dict and can be passed to ``build_posenet`` unchanged.
Flip pairs are what ``DatasetInfo`` (mmpose/datasets/dataset_info.py:90-103) derives from
configs/_base_/datasets/{coco,coco_wholebody}.py.
"""
import copy

_VIT = {
    'small': dict(embed_dim=384, depth=12, num_heads=12, drop_path_rate=0.1),
    'base': dict(embed_dim=768, depth=12, num_heads=12, drop_path_rate=0.3),
    'large': dict(embed_dim=1024, depth=24, num_heads=16, drop_path_rate=0.5),
    'huge': dict(embed_dim=1280, depth=32, num_heads=16, drop_path_rate=0.55),
}

# test_cfg variant (A): 88 of the 115 ViTPose configs — UDP-DARK decode, no shift
TEST_CFG_UDP = dict(flip_test=True, post_process='default', shift_heatmap=False,
                    target_type='GaussianHeatmap', modulate_kernel=11, use_udp=True)
# test_cfg variant (B): 27 configs — quarter-offset decode + 1 px shift of the flipped map
TEST_CFG_SHIFT = dict(flip_test=True, post_process='default', shift_heatmap=True,
                      modulate_kernel=11)

COCO17_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
WHOLEBODY133_FLIP_PAIRS = (
    COCO17_FLIP_PAIRS + [[17, 20], [18, 21], [19, 22]] +
    [[23 + i, 39 - i] for i in range(8)] + [[40 + i, 49 - i] for i in range(5)] +
    [[54, 58], [55, 57], [59, 68], [60, 67], [61, 66], [62, 65], [63, 70], [64, 69],
     [71, 77], [72, 76], [73, 75], [78, 82], [79, 81], [83, 87], [84, 86], [88, 90]] +
    [[91 + i, 112 + i] for i in range(21)])


def flip_pairs_for(num_keypoints):
    if num_keypoints == 17:
        return copy.deepcopy(COCO17_FLIP_PAIRS)
    if num_keypoints == 133:
        return copy.deepcopy(WHOLEBODY133_FLIP_PAIRS)
    # generic: pair (1,2), (3,4) ... like the body part of COCO
    return [[i, i + 1] for i in range(1, num_keypoints - 1, 2)]


def vitpose_model_cfg(size='base', decoder='classic', num_keypoints=17, test_cfg=None):
    """size in {small, base, large, huge}; decoder in {classic, simple}."""
    v = _VIT[size]
    backbone = dict(type='ViT', img_size=(256, 192), patch_size=16, embed_dim=v['embed_dim'],
                    depth=v['depth'], num_heads=v['num_heads'], ratio=1, use_checkpoint=False,
                    mlp_ratio=4, qkv_bias=True, drop_path_rate=v['drop_path_rate'])
    if decoder == 'classic':
        head = dict(type='TopdownHeatmapSimpleHead', in_channels=v['embed_dim'],
                    num_deconv_layers=2, num_deconv_filters=(256, 256), num_deconv_kernels=(4, 4),
                    extra=dict(final_conv_kernel=1), out_channels=num_keypoints,
                    loss_keypoint=dict(type='JointsMSELoss', use_target_weight=True))
    elif decoder == 'simple':
        head = dict(type='TopdownHeatmapSimpleHead', in_channels=v['embed_dim'],
                    num_deconv_layers=0, num_deconv_filters=[], num_deconv_kernels=[], upsample=4,
                    extra=dict(final_conv_kernel=3), out_channels=num_keypoints,
                    loss_keypoint=dict(type='JointsMSELoss', use_target_weight=True))
    else:
        raise ValueError(decoder)
    if test_cfg is None:
        test_cfg = TEST_CFG_SHIFT if num_keypoints == 133 else TEST_CFG_UDP
    return dict(type='TopDown', pretrained=None, backbone=backbone, keypoint_head=head,
                train_cfg=dict(), test_cfg=copy.deepcopy(test_cfg))


def tiny_model_cfg(num_keypoints=5, decoder='classic', test_cfg=None, embed_dim=128, depth=2,
                   num_heads=2, deconv_filters=(64, 64)):
    """A small ViT the reference class accepts as is; used for golden fixtures."""
    cfg = vitpose_model_cfg('base', decoder, num_keypoints, test_cfg)
    cfg['backbone'].update(embed_dim=embed_dim, depth=depth, num_heads=num_heads, drop_path_rate=0.1)
    cfg['keypoint_head']['in_channels'] = embed_dim
    if decoder == 'classic':
        cfg['keypoint_head']['num_deconv_filters'] = tuple(deconv_filters)
    return cfg


# BASELINE.json configs[0..3]
BASELINE_CONFIGS = {
    'S-classic-17': dict(size='small', decoder='classic', num_keypoints=17,
                         test_cfg=dict(flip_test=True, post_process='default',
                                       shift_heatmap=False, modulate_kernel=11, use_udp=False)),
    'B-classic-17': dict(size='base', decoder='classic', num_keypoints=17, test_cfg=TEST_CFG_UDP),
    'L-simple-17': dict(size='large', decoder='simple', num_keypoints=17, test_cfg=TEST_CFG_UDP),
    'H-classic-133': dict(size='huge', decoder='classic', num_keypoints=133,
                          test_cfg=TEST_CFG_SHIFT),
}


def baseline_model_cfg(name):
    return vitpose_model_cfg(**BASELINE_CONFIGS[name])
