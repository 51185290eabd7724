import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import whisper_finetune_b200 as wft
from tests import signals as S
from oracle import pipeline as OP
z = np.load('/root/repo/tests/golden/calculate_mel.npz')
for k in range(int(z['n'])):
    n, n_mels, nv, tp, fp, seed = (int(v) for v in z[f'meta{k}'])
    x = S.make(str(z[f'kind{k}']), n=n, seed=seed)
    fe = wft.FrontEnd(n_mels=n_mels)
    got = fe(x.unsqueeze(0).cuda(), n_valid_frames=None if nv < 0 else [nv], mask_params=None if tp == 0 else z[f'mask{k}'][None, :])[0].cpu()
    want = OP.calculate_mel(x, n_mels, None if nv < 0 else nv, z[f'mask{k}'] if tp > 0 else None)
    d = (got - want).abs()
    print(k, str(z[f'kind{k}']), n, n_mels, nv, z[f'mask{k}'], 'metrics', S.metrics(got, want), 'sub', S.metrics(got[:, ::16], torch.from_numpy(z[f'sub{k}'])), 'argmax', np.unravel_index(int(d.argmax()), d.shape), got.flatten()[d.argmax()].item(), want.flatten()[d.argmax()].item())
