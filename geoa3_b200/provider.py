"""Data formats either side of the hot path: the reference's `.mat` dataset layout on the way
in, its per-instance result files on the way out.

  ModelNet40(data_mat_file, attack_label, resample_num, is_half_forward)
        same constructor / item convention as Provider/modelnet10_instance250.py:14-110
        (file keys `data [M,3,N]`, `normal [M,3,N]`, `label [M,1]`, written by Provider/gen_data_mat.py:297-304)
  write_synthetic_mat(path, instances, npoint)   the synthetic stand-in dataset in that layout
  save_adversarial(saved_dir, name, pc, gt_label, attack_label, est_normal=None)
        the `Mat/<name>.mat` + `PC/<name>.obj` pair of main_attack.py:262-281
  result_name(...)                               the file stem convention of main_attack.py:267

Host-side only (numpy / scipy.io); nothing here touches the GPU."""
import os
from random import choice

import numpy as np
import torch
from scipy.io import loadmat, savemat

from . import synth

ten_label_indexes = list(synth.CLASS_IDS)
ten_label_names = ['airplane', 'bed', 'bookshelf', 'bottle', 'chair', 'monitor', 'sofa', 'table', 'toilet', 'vase']


def _fps_normalized(points, num_points, normal, rng=np.random):
    """Host farthest-point resampling + centre / unit-sphere normalisation (:112-129): random first point,
    Euclidean (not squared) distances, argmax ties -> first."""
    sel = [int(rng.randint(len(points)))]
    dists = np.full(len(points), np.inf)
    for _ in range(num_points - 1):
        dists = np.minimum(dists, np.linalg.norm(points - points[sel[-1]][None, :], axis=1))
        sel.append(int(np.argmax(dists)))
    pts, nrm = np.array(points[sel]), np.array(normal[sel])
    pts = pts - np.average(pts, axis=0)[None, :]
    return pts / np.max(np.linalg.norm(pts, axis=1), axis=0), nrm


class ModelNet40(object):
    """Items (per instance): targeted modes ('All' or a class name) -> [pcs [9,N,3], normals [9,N,3],
    gt_labels [9], target_labels [9]] (the 9 other classes of the ten); 'Untarget' -> [pcs [1,N,3],
    normals [1,N,3], gt_labels [1]]; 'Random' adds one random target of the 40 classes."""

    def __init__(self, data_mat_file='../Data/modelnet10_250instances_1024.mat', attack_label='All', resample_num=-1,
                 is_half_forward=False):
        self.data_root, self.attack_label, self.is_half_forward = data_mat_file, attack_label, is_half_forward
        assert os.path.isfile(data_mat_file), 'No exists .mat file!'
        ds = loadmat(data_mat_file)
        data, normal, label = torch.FloatTensor(ds['data']), torch.FloatTensor(ds['normal']), ds['label']
        if resample_num > 0:
            pairs = [_fps_normalized(data[j].t().numpy(), resample_num, normal[j].t().numpy()) for j in range(data.size(0))]
            data = torch.stack([torch.from_numpy(p).t().float() for p, _ in pairs])
            normal = torch.stack([torch.from_numpy(q).t().float() for _, q in pairs])
        self.start_index = 0
        if attack_label in ten_label_names:   # the file holds 25 instances per class, class-major
            k = ten_label_names.index(attack_label)
            self.start_index = k * 25
            data, normal, label = (x[k * 25:(k + 1) * 25] for x in (data, normal, label))
        else:
            assert attack_label in ('All', 'Untarget', 'Random')
        self.data, self.normal, self.label = data, normal, label

    def __len__(self):
        return self.data.size(0)

    def _views(self, index, copies):
        pc = self.data[index].contiguous().t().unsqueeze(0).expand(copies, -1, -1)
        nr = self.normal[index].contiguous().t().unsqueeze(0).expand(copies, -1, -1)
        return pc, nr

    def __getitem__(self, index):
        label = self.label[index]
        if self.attack_label in ten_label_names or self.attack_label == 'All':
            targets = torch.IntTensor(np.array([i for i in ten_label_indexes if label != i])).long()
            assert targets.size(0) == 9
            gts = torch.IntTensor(label).long().expand_as(targets)
            pcs, nrs = self._views(index, 9)
            if self.is_half_forward:
                return [[pcs[:4], nrs[:4], gts[:4], targets[:4]], [pcs[4:], nrs[4:], gts[4:], targets[4:]]]
            return [pcs, nrs, gts, targets]
        gts = torch.IntTensor(label).long()
        pcs, nrs = self._views(index, 1)
        if self.attack_label == 'Untarget':
            return [pcs, nrs, gts]
        others = [i for i in range(0, 40) if i != gts.item()]
        return [pcs, nrs, gts, torch.IntTensor([choice(others)]).long()]


def write_synthetic_mat(path, instances=250, npoint=1024):
    """25 synthetic instances per class, class-major like the reference file, keys data / normal / label."""
    per = instances // len(ten_label_indexes)
    order = [c + len(ten_label_indexes) * r for c in range(len(ten_label_indexes)) for r in range(per)]
    trip = [synth.make_instance(i, npoint) for i in order]
    savemat(path, {"data": np.stack([t[0] for t in trip]), "normal": np.stack([t[1] for t in trip]),
                   "label": np.asarray([t[2] for t in trip], np.int64)[:, None]})
    return path


def result_name(instance, gt_label, attack_label, expect_label):
    return 'adv_%d_gt%d_attack%d_expect%d' % (instance, gt_label, attack_label, expect_label)


def save_adversarial(saved_dir, name, pc, gt_label, attack_label, est_normal=None):
    """pc [3,N] -> <saved_dir>/Mat/<name>.mat and <saved_dir>/PC/<name>.obj (`v x y z 0 0 0` per point)."""
    for sub in ('Mat', 'PC'):
        os.makedirs(os.path.join(saved_dir, sub), exist_ok=True)
    pc = np.asarray(pc)
    rec = {"adversary_point_clouds": pc, 'gt_label': int(gt_label), 'attack_label': int(attack_label)}
    if est_normal is not None:
        rec['est_normal'] = np.asarray(est_normal)
    savemat(os.path.join(saved_dir, 'Mat', name + '.mat'), rec)
    with open(os.path.join(saved_dir, 'PC', name + '.obj'), 'w') as fout:
        for m in range(pc.shape[1]):
            fout.write('v %f %f %f 0 0 0\n' % (pc[0, m], pc[1, m], pc[2, m]))
