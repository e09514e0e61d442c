import os, sys
sys.path[:0] = ['/root/repo', '/root/repo/speech-separation-project-with-ai_b200']
import numpy as np, torch, sepcore
x = [0.1*torch.randn((1024, 800, 40), device='cuda') for _ in range(3)]; w = 0.05*torch.randn((2,40,129), device='cuda'); b = 0.05*torch.randn((129,), device='cuda')
for _ in range(3): sepcore.conv1d(x[0], w, b, padding='same', activation='sigmoid')
torch.cuda.synchronize()
os.environ['SEPCORE_CONV_TRACE'] = '/tmp/tr.bin'
for rep in range(2):
    sepcore.conv1d(x[1], w, b, padding='same', activation='sigmoid'); torch.cuda.synchronize()
    t=np.fromfile('/tmp/tr.bin',dtype=np.int64).reshape(148,8).astype(np.float64)
    t0=t[:,0].min()
    print('ns since first CTA start: CTA start min/max %.0f/%.0f | prologue done min/mean/max %.0f/%.0f/%.0f | first accumulator ready mean %.0f | last epilogue done min/mean/max %.0f/%.0f/%.0f | CTA end max %.0f' % (
        (t[:,0]-t0).min(), (t[:,0]-t0).max(), (t[:,1]-t0).min(), (t[:,1]-t0).mean(), (t[:,1]-t0).max(), (t[:,2]-t0).mean(), (t[:,3]-t0).min(), (t[:,3]-t0).mean(), (t[:,3]-t0).max(), (t[:,4]-t0).max()))
