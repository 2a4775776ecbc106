import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from neuralnetworklibrary_b200 import testing as syn
from neuralnetworklibrary_b200.retinanet import AnchorGenerator
from neuralnetworklibrary_b200.vision import SSD_loss
from oracle import oracle as orc
dev = torch.device('cuda:0')
H, W, C, B, M = 128, 160, 20, 2, 6
anchors = AnchorGenerator()(torch.zeros(B, 3, H, W, device=dev)); an = orc.anchors(H, W); A = an.shape[0]
gb, gc = syn.make_targets(B, M, H, W, C, seed=3, min_side=12.0, max_frac=0.6)
clas, reg = syn.make_train_activations(B, A, C, seed=3, edge_cases=32)
cd = clas.to(dev).requires_grad_(True); rd = reg.to(dev).requires_grad_(True)
f = SSD_loss(); loss = f([anchors, rd, cd], [gb.to(dev), gc.to(dev)]); loss.backward()
o = orc.loss(an, clas.numpy(), reg.numpy(), gb.numpy(), gc.numpy(), want_matches=True)
for name, got, want in (('dclas', cd.grad.cpu().numpy(), o['dclas']), ('dreg', rd.grad.cpu().numpy(), o['dreg'])):
    nz = want != 0
    rel = np.zeros_like(want); rel[nz] = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    k = np.unravel_index(rel.argmax(), rel.shape)
    print(name, 'max rel', rel.max(), 'at', k, 'got', got[k], 'want', want[k], 'scale max', np.abs(want).max(), 'n>1e-5:', (rel > 1e-5).sum())
    if name == 'dclas':
        print('  x =', clas.numpy()[k], 'match', o['matches'][k[0], k[1]])
