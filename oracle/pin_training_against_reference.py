"""TEST INFRASTRUCTURE ONLY -- golden vectors for the TRAINING step, generated from the unmodified reference.

Imports /root/reference (same shims as pin_against_reference.py: `.cuda()` -> identity, the CUDA op replaced by the oracle's
differentiable restatement), runs the reference's own training forward (models/mode_disparity.py:99-155, three heads, batch-stat
BatchNorm) and loss (train_disparity.py:147-158: 0.5 / 0.7 / 1.0 smooth-L1 on masked pixels) on a seeded tiny configuration,
back-propagates -- once in fp32 and once in fp64 -- and stores predictions, loss, a selection of fp64 parameter gradients and
the reference's own fp32-vs-fp64 distance for each in
tests/golden/mode_disparity_train_tiny_cassini.npz.  tests/test_gpu_model.py compares the product's training path
(libmode_b200 spherical forward + backward kernels under autograd) against it.  Cannot travel to the GPU box; run here:

    python oracle/pin_training_against_reference.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pin_against_reference as P  # noqa: E402
from tests import helpers as Hh  # noqa: E402

GRAD_KEYS, train_inputs, loss_fn = Hh.TRAIN_GRAD_KEYS, Hh.train_inputs, Hh.train_loss


def run_reference(models, dt, name='tiny_cassini'):
  """One training step of the unmodified reference in dtype dt (fp64 = the truth two fp32 evaluations are measured against)."""
  sd, (H, W, D, st, seed), _ = Hh.golden_state_dict(name)
  left, right, disp_true, mask = train_inputs(H, W, D, seed)  # drawn in fp32 in both runs
  torch.set_default_dtype(dt)
  float_tensor = torch.FloatTensor
  torch.FloatTensor = torch.DoubleTensor if dt == torch.float64 else float_tensor  # the reference allocates its cost volume with it
  try:
    model = models.ModeDisparity(D, conv='Sphere', in_height=H, in_width=W, sphereType=st, out_conf=False)
    model.load_state_dict(sd)
    model = model.to(dt).train()
    for m in model.modules():
      if hasattr(m, 'position'):
        m.position = m.position.to(dt)
    o1, o2, o3 = model(left.to(dt), right.to(dt))
    loss = loss_fn(o1, o2, o3, disp_true.to(dt), mask)
    loss.backward()
  finally:
    torch.set_default_dtype(torch.float32)
    torch.FloatTensor = float_tensor
  grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k in GRAD_KEYS}
  rm = dict(model.named_buffers())['dres0.0.1.running_mean'].detach().clone()
  return (o1.detach(), o2.detach(), o3.detach()), loss.item(), grads, rm


def main():
  models, SCM, RG = P.import_reference()
  torch.set_num_threads(os.cpu_count())
  name = sys.argv[1] if len(sys.argv) > 1 else 'tiny_cassini'
  (o1, o2, o3), loss, g32, rm = run_reference(models, torch.float32, name)
  (p1, p2, p3), loss64, g64, _ = run_reference(models, torch.float64, name)
  out = dict(pred1=o1.numpy(), pred2=o2.numpy(), pred3=o3.numpy(), loss=np.float64(loss), loss64=np.float64(loss64))
  out['running_mean/dres0.0.1'] = rm.numpy()  # batch-stat BN also updates the running statistics: pin one of them
  print('loss fp32 %.6f fp64 %.6f   pred3 range [%.3f, %.3f]' % (loss, loss64, o3.min().item(), o3.max().item()))
  for k in GRAD_KEYS:
    # the fixture holds the fp64 gradient and how far the reference's OWN fp32 evaluation is from it: batch-statistics
    # BatchNorm over 16 samples per channel at 1/16 resolution makes the step ill-conditioned, the test scales its bound by it
    a, b = Hh.grad_sample(g32[k]).double(), Hh.grad_sample(g64[k])
    floor = ((a - b).norm() / b.norm()).item()
    out['grad64/' + k] = b.numpy().astype(np.float32)
    out['floor/' + k] = np.float64(floor)
    print('%-55s |g|max %.3e   reference fp32 vs fp64 %.2e' % (k, g64[k].abs().max().item(), floor))
  np.savez_compressed(os.path.join(P.GOLD, f'mode_disparity_train_{name}.npz'), **out)


if __name__ == '__main__':
  main()
