"""Which stored tensors of the 3-D stack cost the end-point error when they are kept in 16 bits?  (CPU, fp32 arithmetic.)

An fp32 restatement of the 3-D stack (oracle.regularise) with a hook that rounds selected *stored* activations to bf16 / fp16,
exactly where the tensor-core plan stores them (every conv output after BN/residual/ReLU; logits stay fp32).  Prints the EPE of the
soft-argmin disparity against the un-rounded stack for several storage policies.  VERDICT r01 task 1.

    python tools/precision_study.py [tiny_cassini|small_cassini]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mode_oracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402

TRUNK = {'c0', 'out1', 'out2', 'out3'}            # cost0 and the hourglass outputs (out_k + cost0)
SKIPS = {'pre1', 'post1', 'pre2', 'post2', 'pre3', 'post3'}


def stack(m, cost, q):
  """regularise() with q(name, tensor) applied to every tensor the 16-bit plan stores."""
  cb, db = m.convbn3d, m.deconvbn3d
  c0 = q('dres0.0', F.relu(cb(cost, 'dres0.0', 1)))
  c0 = q('dres0.2', F.relu(cb(c0, 'dres0.2', 1)))
  t = q('dres1.0', F.relu(cb(c0, 'dres1.0', 1)))
  c0 = q('c0', cb(t, 'dres1.2', 1) + c0)

  def hg(i, key, x, presqu, postsqu):
    out = q(f'hg{i}.c1', F.relu(cb(x, key + '.conv1.0', 2)))
    pre = cb(out, key + '.conv2', 1)
    pre = q(f'pre{i}', F.relu(pre + postsqu) if postsqu is not None else F.relu(pre))
    out = q(f'hg{i}.c3', F.relu(cb(pre, key + '.conv3.0', 2)))
    out = q(f'hg{i}.c4', F.relu(cb(out, key + '.conv4.0', 1)))
    post = q(f'post{i}', F.relu(db(out, key + '.conv5') + (presqu if presqu is not None else pre)))
    return q(f'out{i}', db(post, key + '.conv6') + c0), pre, post

  out1, pre1, post1 = hg(1, 'dres2', c0, None, None)
  out2, pre2, post2 = hg(2, 'dres3', out1, pre1, post1)
  out3, pre3, post3 = hg(3, 'dres4', out2, pre1, post2)

  def classif(i, x, key):
    x = q(f'cls{i}', F.relu(cb(x, key + '.0', 1)))
    return F.conv3d(x, m.sd[key + '.2.weight'], None, 1, 1)

  cost1 = classif(1, out1, 'classif1')
  cost2 = classif(2, out2, 'classif2') + cost1
  return classif(3, out3, 'classif3') + cost2


def main():
  name = sys.argv[1] if len(sys.argv) > 1 else 'tiny_cassini'
  sd, (H, W, D, st, seed), z = Hh.golden_state_dict(name)
  left, right = Hh.synth_inputs(H, W, seed)
  m = O._SD(sd)
  pos = torch.from_numpy(O.gen_sphere_position(H // 4, W // 4, st))
  with torch.no_grad():
    fl, fr = O.feature_extraction(m, left, pos), O.feature_extraction(m, right, pos)
    cost = O.cost_volume(fl, fr, D // 4)
    ref = O.disparity_regression(stack(m, cost, lambda n, t: t), D, H, W)

    def run(policy):
      def q(n, t):
        dt = policy(n)
        return t if dt is None else t.to(dt).float()
      c = cost.to(policy('cost') or torch.float32).float()
      return (O.disparity_regression(stack(m, c, q), D, H, W) - ref).abs().mean().item()

    bf, hf = torch.bfloat16, torch.float16
    print(f'{name}: EPE of the soft-argmin disparity vs the un-rounded fp32 stack (px)')
    print('  all bf16                         %.5f' % run(lambda n: bf))
    print('  all fp16                         %.5f' % run(lambda n: hf))
    print('  bf16, trunk fp32                 %.5f' % run(lambda n: None if n in TRUNK else bf))
    print('  bf16, trunk+skips fp32           %.5f' % run(lambda n: None if n in TRUNK | SKIPS else bf))
    print('  bf16, trunk+skips fp16           %.5f' % run(lambda n: hf if n in TRUNK | SKIPS else bf))
    print('  bf16, trunk+skips+cls fp32       %.5f' % run(lambda n: None if (n in TRUNK | SKIPS or n.startswith('cls')) else bf))
    print('  bf16 only cls (rest fp32)        %.5f' % run(lambda n: bf if n.startswith('cls') else None))
    print('  bf16 only cost+dres (rest fp32)  %.5f' % run(lambda n: bf if (n == 'cost' or n.startswith('dres')) else None))
    print('  bf16 only hg interiors           %.5f' % run(lambda n: bf if n.startswith('hg') else None))
    print('  bf16 only trunk+skips            %.5f' % run(lambda n: bf if n in TRUNK | SKIPS else None))


if __name__ == '__main__':
  main()
