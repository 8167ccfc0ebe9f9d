import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import torch, _native as nv, sif_functions as sf
torch.manual_seed(0)
X = (0.4 * torch.randn(50000, 300, device='cuda') + 0.3 * torch.randn(1, 300, device='cuda')).contiguous()
G = sf.gram(X, nv.GRAM_TF32X3)
torch.save(G.cpu(), '/tmp/g_%s_%s.pt' % (os.environ.get('MMB_TC_NOHI', 'hi'), os.environ.get('MMB_TC_PASSES', '3')))
want = X.double().T @ X.double()
print(os.environ.get('MMB_TC_NOHI'), os.environ.get('MMB_TC_PASSES'), 'err', ((G.double() - want).abs().max() / want.abs().max()).item())
