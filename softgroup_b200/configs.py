"""Model hyper-parameters of the reference configs that BASELINE.json's workloads are quoted on
(values restated from /root/reference/configs/softgroup/*.yaml; data, not code).
`model_cfg(name)` returns kwargs for softgroup_b200.model.SoftGroup."""
import copy

_SCANNET = dict(  # configs/softgroup/softgroup_scannet.yaml:1-29
    channels=32, num_blocks=7, semantic_classes=20, instance_classes=18, sem2ins_classes=[], semantic_only=False,
    ignore_label=-100,
    grouping_cfg=dict(score_thr=0.2, radius=0.04, mean_active=300,
                      class_numpoint_mean=[-1., -1., 3917., 12056., 2303., 8331., 3948., 3166., 5629., 11719., 1003.,
                                           3317., 4912., 10221., 3889., 4136., 2120., 945., 3967., 2589.],
                      npoint_thr=0.05, ignore_classes=[0, 1]),
    instance_voxel_cfg=dict(scale=50, spatial_shape=20),
    train_cfg=dict(max_proposal_num=200, pos_iou_thr=0.5),
    test_cfg=dict(x4_split=False, cls_score_thr=0.001, mask_score_thr=-0.5, min_npoint=100,
                  eval_tasks=['semantic', 'instance']),
    fixed_modules=[])

_S3DIS = dict(  # configs/softgroup/softgroup_s3dis_fold5.yaml:1-28
    channels=32, num_blocks=7, semantic_classes=13, instance_classes=13, sem2ins_classes=[0, 1], semantic_only=False,
    ignore_label=-100,
    grouping_cfg=dict(score_thr=0.2, radius=0.04, mean_active=300,
                      class_numpoint_mean=[34229, 39796, 12210, 7457, 5439, 10225, 6016, 1724, 5092, 7424, 5279, 6189,
                                           1823],
                      npoint_thr=0.05, ignore_classes=[0, 1]),
    instance_voxel_cfg=dict(scale=50, spatial_shape=20),
    train_cfg=dict(max_proposal_num=200, pos_iou_thr=0.5),
    test_cfg=dict(x4_split=True, cls_score_thr=0.001, mask_score_thr=-0.5, min_npoint=100,
                  eval_tasks=['semantic', 'instance']),
    fixed_modules=[])

_KITTI = dict(  # configs/softgroup/softgroup_kitti.yaml:1-31
    in_channels=1, channels=32, num_blocks=7, semantic_classes=19, instance_classes=8, sem2ins_classes=[],
    semantic_only=False, ignore_label=-100, with_coords=False,
    grouping_cfg=dict(score_thr=0.2, radius=0.1, mean_active=300, class_numpoint_mean=[-1.] * 19, npoint_thr=5,
                      ignore_classes=list(range(11))),
    instance_voxel_cfg=dict(scale=20, spatial_shape=20),
    train_cfg=dict(max_proposal_num=200, pos_iou_thr=0.5),
    test_cfg=dict(x4_split=False, cls_score_thr=0.1, mask_score_thr=-0.5, min_npoint=25, eval_tasks=['panoptic'],
                  panoptic_skip_iou=0.5),
    fixed_modules=[])

CONFIGS = {'scannet': _SCANNET, 's3dis': _S3DIS, 'kitti': _KITTI}
# which synthetic shape (softgroup_b200.synth.SHAPES) goes with which config
SHAPE_OF = {'scannet': 'c2_scannet', 's3dis': 'c3_s3dis', 'kitti': 'c4_kitti'}


def model_cfg(name='scannet', **overrides):
    cfg = copy.deepcopy(CONFIGS[name])
    for k, v in overrides.items():
        if isinstance(v, dict) and isinstance(cfg.get(k), dict):
            cfg[k].update(v)
        else:
            cfg[k] = v
    return cfg
