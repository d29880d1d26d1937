"""TEST INFRASTRUCTURE — generates tests/golden/preprocess.npz with the reference's own functions (get_warp_matrix,
get_affine_transform loaded by path from /root/reference) + real cv2.warpAffine + torch to_tensor/normalize maths.
Run in the authoring container:  python -m oracle.make_golden_preprocess"""
import os

import cv2
import numpy as np
import torch

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'preprocess.npz')


def main():
    ref_loader.load_reference()
    import sys
    post = sys.modules['mmpose.core.post_processing']
    rng = np.random.RandomState(7)
    small = rng.randint(0, 256, size=(15, 20, 3)).astype(np.uint8)   # blocky image: compresses, still exercises every weight
    img = cv2.resize(small, (160, 120), interpolation=cv2.INTER_CUBIC)
    img[::7, ::5] = rng.randint(0, 256, size=img[::7, ::5].shape).astype(np.uint8)
    boxes = np.array([[20, 10, 60, 90], [-10, 30, 80, 70], [70, 60, 9, 30]],
                     dtype=np.float32)
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    out = dict(img=img, boxes=boxes)
    for use_udp in (True, False):
        crops, mats, cs = [], [], []
        for b in boxes:
            x, y, w, h = b
            ar = 192 / 256
            center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
            if w > ar * h:
                h = w * 1.0 / ar
            elif w < ar * h:
                w = h * ar
            scale = np.array([w / 200.0, h / 200.0], dtype=np.float32) * 1.25
            size = np.array([192, 256])
            if use_udp:
                trans = post.get_warp_matrix(0, center * 2.0, size - 1.0, scale * 200.0)
            else:
                trans = post.get_affine_transform(center, scale, 0, size)
            crop = cv2.warpAffine(img, trans, (192, 256), flags=cv2.INTER_LINEAR)
            t = torch.from_numpy(crop).permute(2, 0, 1).float().div(255)
            t = t.sub(torch.from_numpy(mean)[:, None, None]).div(torch.from_numpy(std)[:, None, None])
            crops.append(t.numpy().astype(np.float16 if False else np.float32))
            mats.append(np.asarray(trans, dtype=np.float64))
            cs.append(np.concatenate([center, scale]))
        tag = 'udp' if use_udp else 'affine'
        # store the uint8 warps (small) and one full float crop; the float maths is checked separately
        out[f'{tag}_mats'] = np.stack(mats)
        out[f'{tag}_cs'] = np.stack(cs)
        out[f'{tag}_crop0_f32'] = crops[0]
        out[f'{tag}_warps_u8'] = np.stack([cv2.warpAffine(img, m.astype(np.float32) if use_udp else m, (192, 256),
                                                         flags=cv2.INTER_LINEAR) for m in mats])
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == '__main__':
    main()
