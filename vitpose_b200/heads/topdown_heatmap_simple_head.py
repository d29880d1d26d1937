"""``TopdownHeatmapSimpleHead`` with the reference's registry name, constructor signature, error behaviour and
state-dict layout (mmpose/models/heads/topdown_heatmap_simple_head.py:16-350 and
topdown_heatmap_base_head.py:40-120), executed by the sm_100a implicit-GEMM kernels.

Supported decoders are the two the ViTPose configs use: "classic" (k4/s2/p1 deconvs + BN + ReLU, 1x1 final
conv) and "simple" (ReLU + bilinear x``upsample`` + 3x3 final conv).  Parameters live in the same
``deconv_layers.{0,1,3,4}`` / ``final_layer`` slots as the reference so checkpoints load unchanged; the
modules are parameter holders and are never called.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops
from ..builder import HEADS, build_loss
from ..core.post_processing import flip_index_from_pairs
from ..engine import BF16, decode_mode_from_cfg, fold_bn, pack_deconv_weight


class TopdownHeatmapBaseHead(nn.Module):
    """Base class: ``decode`` (topdown_heatmap_base_head.py:40-103) and ``_get_deconv_cfg`` (:105-120)."""

    def decode(self, img_metas, output, **kwargs):
        """img_metas: list of dicts (center, scale, image_file, [bbox_score], [bbox_id]);
        output: np.ndarray[N,K,H,W] heatmaps.  Returns the reference's result dict."""
        from ..core.evaluation import keypoints_from_heatmaps
        batch_size = len(img_metas)
        bbox_ids = [] if 'bbox_id' in img_metas[0] else None
        c = np.zeros((batch_size, 2), dtype=np.float32)
        s = np.zeros((batch_size, 2), dtype=np.float32)
        image_paths = []
        score = np.ones(batch_size)
        for i in range(batch_size):
            c[i, :] = img_metas[i]['center']
            s[i, :] = img_metas[i]['scale']
            image_paths.append(img_metas[i]['image_file'])
            if 'bbox_score' in img_metas[i]:
                score[i] = np.array(img_metas[i]['bbox_score']).reshape(-1)[0]
            if bbox_ids is not None:
                bbox_ids.append(img_metas[i]['bbox_id'])
        preds, maxvals = keypoints_from_heatmaps(
            output, c, s,
            unbiased=self.test_cfg.get('unbiased_decoding', False),
            post_process=self.test_cfg.get('post_process', 'default'),
            kernel=self.test_cfg.get('modulate_kernel', 11),
            valid_radius_factor=self.test_cfg.get('valid_radius_factor', 0.0546875),
            use_udp=self.test_cfg.get('use_udp', False),
            target_type=self.test_cfg.get('target_type', 'GaussianHeatmap'))
        return pack_results(preds, maxvals, c, s, score, image_paths, bbox_ids)

    @staticmethod
    def _get_deconv_cfg(deconv_kernel):
        if deconv_kernel == 4:
            padding, output_padding = 1, 0
        elif deconv_kernel == 3:
            padding, output_padding = 1, 1
        elif deconv_kernel == 2:
            padding, output_padding = 0, 0
        else:
            raise ValueError(f'Not supported num_kernels ({deconv_kernel}).')
        return deconv_kernel, padding, output_padding


def pack_results(preds, maxvals, c, s, score, image_paths, bbox_ids):
    """all_preds / all_boxes packing of topdown_heatmap_base_head.py:87-103."""
    n = preds.shape[0]
    all_preds = np.zeros((n, preds.shape[1], 3), dtype=np.float32)
    all_boxes = np.zeros((n, 6), dtype=np.float32)
    all_preds[:, :, 0:2] = preds[:, :, 0:2]
    all_preds[:, :, 2:3] = maxvals
    all_boxes[:, 0:2] = c[:, 0:2]
    all_boxes[:, 2:4] = s[:, 0:2]
    all_boxes[:, 4] = np.prod(s * 200.0, axis=1)
    all_boxes[:, 5] = score
    return dict(preds=all_preds, boxes=all_boxes, image_paths=image_paths, bbox_ids=bbox_ids)


@HEADS.register_module()
class TopdownHeatmapSimpleHead(TopdownHeatmapBaseHead):

    def __init__(self,
                 in_channels,
                 out_channels,
                 num_deconv_layers=3,
                 num_deconv_filters=(256, 256, 256),
                 num_deconv_kernels=(4, 4, 4),
                 extra=None,
                 in_index=0,
                 input_transform=None,
                 align_corners=False,
                 loss_keypoint=None,
                 train_cfg=None,
                 test_cfg=None,
                 upsample=0,):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        if loss_keypoint is None:
            raise TypeError('loss_keypoint config is required (the reference fails in build_loss(None))')
        self.loss = build_loss(loss_keypoint)
        self.upsample = upsample
        self.train_cfg = {} if train_cfg is None else train_cfg
        self.test_cfg = {} if test_cfg is None else test_cfg
        self.target_type = self.test_cfg.get('target_type', 'GaussianHeatmap')

        self._init_inputs(in_channels, in_index, input_transform)
        self.in_index = in_index
        self.align_corners = align_corners
        self._head_in_channels = self.in_channels

        if extra is not None and not isinstance(extra, dict):
            raise TypeError('extra should be dict or None.')
        self.extra = extra

        self.num_deconv_layers = num_deconv_layers
        self.num_deconv_filters = tuple(num_deconv_filters)[:max(num_deconv_layers, 0)]
        self.num_deconv_kernels = tuple(num_deconv_kernels)[:max(num_deconv_layers, 0)]
        if num_deconv_layers > 0:
            self.deconv_layers = self._make_deconv_layer(num_deconv_layers, num_deconv_filters, num_deconv_kernels)
        elif num_deconv_layers == 0:
            self.deconv_layers = nn.Identity()
        else:
            raise ValueError(f'num_deconv_layers ({num_deconv_layers}) should >= 0.')

        identity_final_layer = False
        if extra is not None and 'final_conv_kernel' in extra:
            assert extra['final_conv_kernel'] in [0, 1, 3]
            if extra['final_conv_kernel'] == 3:
                padding = 1
            elif extra['final_conv_kernel'] == 1:
                padding = 0
            else:
                identity_final_layer = True
            kernel_size = extra['final_conv_kernel']
        else:
            kernel_size, padding = 1, 0
        self.final_conv_kernel = kernel_size

        if identity_final_layer:
            self.final_layer = nn.Identity()
        else:
            conv_channels = num_deconv_filters[-1] if num_deconv_layers > 0 else self.in_channels
            if extra is not None and extra.get('num_conv_layers', 0) > 0:
                raise NotImplementedError("extra['num_conv_layers'] > 0 is not used by any ViTPose config")
            self.final_layer = nn.Conv2d(conv_channels, out_channels, kernel_size, stride=1, padding=padding)
        self._packed = None
        self._packed_key = None

    # ---- construction helpers (same checks / errors as the reference) --------------------------------
    def _init_inputs(self, in_channels, in_index, input_transform):
        if input_transform is not None:
            assert input_transform in ['resize_concat', 'multiple_select']
            raise NotImplementedError('multi-level input transforms are not used by any ViTPose config')
        self.input_transform = input_transform
        assert isinstance(in_channels, int)
        assert isinstance(in_index, int)
        self.in_channels = in_channels

    def _make_deconv_layer(self, num_layers, num_filters, num_kernels):
        if num_layers != len(num_filters):
            raise ValueError(f'num_layers({num_layers}) != length of num_filters({len(num_filters)})')
        if num_layers != len(num_kernels):
            raise ValueError(f'num_layers({num_layers}) != length of num_kernels({len(num_kernels)})')
        layers = []
        for i in range(num_layers):
            kernel, padding, output_padding = self._get_deconv_cfg(num_kernels[i])
            planes = num_filters[i]
            layers.append(nn.ConvTranspose2d(self.in_channels, planes, kernel_size=kernel, stride=2,
                                             padding=padding, output_padding=output_padding, bias=False))
            layers.append(nn.BatchNorm2d(planes))
            layers.append(nn.ReLU(inplace=True))
            self.in_channels = planes
        return nn.Sequential(*layers)

    def init_weights(self):
        """simple_head.py:339-350: N(0, 0.001) for deconv / final conv, BN = 1."""
        for m in self.modules():
            if isinstance(m, (nn.ConvTranspose2d, nn.Conv2d)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._packed = None

    def cfg_dict(self):
        return dict(in_channels=self._head_in_channels, out_channels=self.out_channels,
                    num_deconv_layers=self.num_deconv_layers, num_deconv_filters=self.num_deconv_filters,
                    num_deconv_kernels=self.num_deconv_kernels, upsample=self.upsample,
                    extra=dict(final_conv_kernel=self.final_conv_kernel))

    def _weights_version(self):
        vals = list(self.parameters()) + list(self.buffers())
        return tuple((p.data_ptr(), p._version) for p in vals)

    # ---- loss / accuracy (training config; evaluated with the registered loss module) -----------------
    def get_loss(self, output, target, target_weight):
        losses = dict()
        assert not isinstance(self.loss, nn.Sequential)
        assert target.dim() == 4 and target_weight.dim() == 3
        losses['heatmap_loss'] = self.loss(output, target, target_weight)
        return losses

    def get_accuracy(self, output, target, target_weight):
        """simple_head.py:170-195. The reference copies both heatmap batches to the host every iteration and returns
        a Python float; here arg-max and the PCK arithmetic stay on the device and ``acc_pose`` is a 0-dim CUDA tensor
        (``_parse_losses`` and loggers call ``.mean()`` / ``.item()`` on it when they need the number)."""
        from .. import ops
        accuracy = dict()
        if self.target_type == 'GaussianHeatmap':
            _, avg_acc, _ = ops.pose_pck_accuracy(output.detach().float().contiguous(), target.float().contiguous(),
                                                  target_weight.detach().squeeze(-1))
            accuracy['acc_pose'] = avg_acc[0]
        return accuracy

    # ---- execution --------------------------------------------------------------------------------------
    def _packed_weights(self, device):
        key = (self._weights_version(), str(device))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        pk = dict(deconv=[])
        for i in range(self.num_deconv_layers):
            conv, bn = self.deconv_layers[3 * i], self.deconv_layers[3 * i + 1]
            if conv.kernel_size != (4, 4):
                raise NotImplementedError('only kernel-4 deconvs are implemented (all ViTPose configs)')
            scale, shift = fold_bn(bn.weight.detach().float(), bn.bias.detach().float(),
                                   bn.running_mean.float(), bn.running_var.float(), bn.eps)
            pk['deconv'].append((pack_deconv_weight(conv.weight.detach().float().to(device)),
                                 scale.to(device), shift.to(device)))
        if isinstance(self.final_layer, nn.Conv2d):
            fw = self.final_layer.weight.detach().float()
            K, cin, kh, kw = fw.shape
            fw = fw.reshape(K, cin) if kh == 1 else fw.permute(0, 2, 3, 1).reshape(K, kh * kw * cin)
            pk['final_w'] = fw.to(device).to(BF16).contiguous()
            pk['final_b'] = self.final_layer.bias.detach().float().to(device).contiguous()
        self._packed, self._packed_key = pk, key
        return pk

    def forward(self, x):
        """x: features [N, C, h, w] (fp32 CUDA, NCHW as the reference backbone returns) -> heatmaps fp32 NCHW."""
        if isinstance(x, (list, tuple)):
            x = x[self.in_index]
        if not x.is_cuda:
            raise _lib.VitposeLibError('vitpose_b200 has no CPU path: features must be CUDA tensors')
        pk = self._packed_weights(x.device)
        nhwc = x.permute(0, 2, 3, 1).contiguous().to(BF16)      # layout/dtype plumbing for standalone calls
        if self.upsample > 0:
            nhwc = ops.relu_upsample_nhwc(nhwc, self.upsample)
        for w, scale, shift in pk['deconv']:
            nhwc = ops.deconv4x4s2_bn_relu(nhwc, w, scale, shift)
        n, h, w_, c = nhwc.shape
        if not isinstance(self.final_layer, nn.Conv2d):
            return nhwc.permute(0, 3, 1, 2).float().contiguous()
        if self.final_conv_kernel == 1:
            out = ops.gemm(nhwc.reshape(n * h * w_, c), pk['final_w'], _lib.EPI_NCHW_F32, bias=pk['final_b'],
                           period=h * w_)
            return out.reshape(n, -1, h, w_)
        return ops.conv3x3_nchw(nhwc, pk['final_w'], pk['final_b'])

    def inference_model(self, x, flip_pairs=None):
        """Returns np.ndarray heatmaps; with flip_pairs the maps are flipped back (and shifted when
        test_cfg['shift_heatmap']), as simple_head.py:204-227."""
        output = self.forward(x)
        if flip_pairs is not None:
            perm = torch.from_numpy(flip_index_from_pairs(output.shape[1], flip_pairs)).to(output.device)
            output = ops.flip_back(output.contiguous(), perm, self.test_cfg.get('shift_heatmap', False))
        return output.detach().cpu().numpy()
