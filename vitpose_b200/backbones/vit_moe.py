"""``ViTMoE`` — the ViTPose+ backbone (mmpose/models/backbones/vit_moe.py:77-115 MoEMlp, :218-380 ViTMoE) with the
reference's registry name, constructor and state-dict layout.

In ViTPose+ every block's FFN output is ``cat(fc2(h), experts[d](h))``: the first ``D - part_features`` channels
come from a shared ``fc2`` [D - part, 4D], the last ``part_features`` from the expert of the crop's dataset ``d``.
For one dataset that is an ordinary Linear whose weight is ``cat(fc2.weight, experts[d].weight)`` — exactly what the
reference's ``tools/model_split.py:36-40`` writes out per dataset — so the B200 path is the ViT engine run on that
effective weight, one packed engine per dataset index (built lazily, cached). A batch that mixes datasets is
grouped by index by the caller (``TopDownMoE.forward_test``).
"""
import torch
import torch.nn as nn

from ..builder import BACKBONES
from ..engine import VitPoseEngine
from .vit import ViT, _ParamHolder


class _MoEMlp(_ParamHolder):
    def __init__(self, num_expert, dim, hidden, part_features):
        super().__init__()
        self.part_features = part_features
        self.num_expert = num_expert
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim - part_features)
        self.experts = nn.ModuleList([nn.Linear(hidden, part_features) for _ in range(num_expert)])


@BACKBONES.register_module()
class ViTMoE(ViT):

    def __init__(self, *args, num_expert=1, part_features=None, **kwargs):
        super().__init__(*args, **kwargs)
        if part_features is None:
            raise TypeError('ViTMoE needs part_features (channels produced by the dataset experts)')
        self.num_expert = num_expert
        self.part_features = part_features
        D = self.embed_dim
        hidden = int(D * self._cfg['mlp_ratio'])
        for blk in self.blocks:
            blk.mlp = _MoEMlp(num_expert, D, hidden, part_features)
        self._engines = {}
        self._freeze_stages()            # the FFN modules were replaced after ViT.__init__ froze the old ones
        self.init_weights(None)

    def effective_state_dict(self, dataset_idx):
        """State dict of the plain ViT that this backbone is for crops of dataset ``dataset_idx``
        (tools/model_split.py:36-40,70-79: fc2 <- cat(fc2, experts[d]) on the output dimension)."""
        if not 0 <= dataset_idx < self.num_expert:
            raise IndexError(f'dataset_idx {dataset_idx} outside the {self.num_expert} experts')
        out = {}
        for k, v in self.state_dict().items():
            if '.mlp.experts.' in k:
                continue
            if '.mlp.fc2.' in k:
                e = self.state_dict()[k.replace('fc2.', f'experts.{dataset_idx}.')]
                v = torch.cat([v, e], dim=0)
            out[k] = v
        return out

    def engine(self, head=None, dataset_idx=0):
        key = (dataset_idx, self._weights_version(), None if head is None else head._weights_version())
        cached = self._engines.get(dataset_idx)
        if cached is None or cached[0] != key:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('vitpose_b200 has no CPU path: move the model to a CUDA device (model.cuda())')
            sd = {'backbone.' + k: v for k, v in self.effective_state_dict(dataset_idx).items()}
            head_cfg = None
            if head is not None:
                sd.update({'keypoint_head.' + k: v for k, v in head.state_dict().items()})
                head_cfg = head.cfg_dict()
            # The per-dataset networks differ in mlp.fc2 only (vit_moe.py:107-111: the expert owns the last
            # part_features output columns): every other repacked tensor is shared with an engine already built for
            # another dataset at the same weights version, so num_expert engines cost one backbone + num_expert x fc2.
            donor = next((e for k, e in self._engines.values() if k[1:] == key[1:] and e.device == dev), None)
            cached = (key, VitPoseEngine(self._cfg, head_cfg, sd, device=dev, share_weights_from=donor,
                                         private_keys=('.mlp.fc2.',)))
            self._engines[dataset_idx] = cached
        return cached[1]

    def moe_engine(self, head=None):
        """ONE engine for batches that mix datasets: the packed network of dataset 0 plus every dataset's mlp.fc2
        (cat(fc2, experts[d]), vit_moe.py:107-111) on the device; ``VitPoseEngine.set_moe_runs`` then selects the fc2
        per run of crops inside a single forward pass (include/vitpose_b200.h: vpb_moe_runs)."""
        eng = self.engine(head, 0)
        key = (self._weights_version(), None if head is None else head._weights_version())
        if getattr(eng, '_experts_key', None) != key:
            dev, sd = eng.device, self.state_dict()
            experts = {}
            for d in range(self.num_expert):
                ws_, bs_ = [], []
                for i in range(len(self.blocks)):
                    pre = f'blocks.{i}.mlp.'
                    w = torch.cat([sd[pre + 'fc2.weight'], sd[pre + f'experts.{d}.weight']], dim=0)
                    b = torch.cat([sd[pre + 'fc2.bias'], sd[pre + f'experts.{d}.bias']], dim=0)
                    ws_.append(w.detach().to(device=dev, dtype=torch.float32).to(torch.bfloat16).contiguous())
                    bs_.append(b.detach().to(device=dev, dtype=torch.float32).contiguous())
                experts[d] = (ws_, bs_)
            eng.set_experts(experts)
            eng._experts_key = key
        return eng

    @staticmethod
    def dataset_runs(src):
        """(order, runs): a stable permutation that sorts the crops by dataset index and the [(dataset, count)] runs."""
        src = [int(v) for v in src]
        order = sorted(range(len(src)), key=lambda i: src[i])
        runs = []
        for i in order:
            if runs and runs[-1][0] == src[i]:
                runs[-1][1] += 1
            else:
                runs.append([src[i], 1])
        return order, [(d, c) for d, c in runs]

    def forward_features(self, x, dataset_source=None):
        """[N,3,H,W] -> [N,D,Hp,Wp]; ``dataset_source`` int tensor [N] (default: dataset 0 for every crop). Mixed batches
        run in one pass: crops sorted by dataset, one mlp.fc2 launch per run."""
        from .. import ops
        n = x.shape[0]
        src = [0] * n if dataset_source is None else dataset_source.detach().cpu().long().tolist()
        order, runs = self.dataset_runs(src)
        eng = self.moe_engine(None)
        perm = torch.tensor(order, device=x.device)
        eng.set_moe_runs(runs)
        try:
            _, tokens = eng.forward_heatmaps(x.index_select(0, perm).float(), flip=False, want_features=True,
                                             want_heatmaps=False)
        finally:
            eng.set_moe_runs(None)
        hp, wp = eng.tokens_hw
        f = ops.tokens_to_nchw(tokens, hp, wp)
        out = torch.empty_like(f)
        out.index_copy_(0, perm, f)
        return out

    def forward(self, x, dataset_source=None):
        return self.forward_features(x, dataset_source)
