import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from tests.util_golden import token_cases, pack_tokens
from pybpl_b200 import ops
d,toks=token_cases()
DEV='cuda:0'
T=lambda a,dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a),dtype=dt,device=DEV)
def run(sel):
    b=pack_tokens(sel)
    aff=b['affine'].copy(); aff[~b['has_affine']]=[1,1,0,0]
    use=bool(b['has_affine'].any())
    batch=ops.TokenBatch(b['tok_stroke_off'],b['stroke_sub_off'],T(b['ctrl']),T(b['invscale']),T(b['position']),T(b['eps']),T(b['sigma']),affine=T(aff) if use else None)
    imgs=ops.pack_images(T(b['images'],torch.uint8))
    ll,off=ops.score_tokens(batch,imgs,T(b['img_idx'],torch.int32))
    return ll.cpu().numpy()
ll=run(toks)
ref=np.array([t['ll'] for t in toks])
rel=np.abs(ll-ref)/np.abs(ref)
bad=np.nonzero(rel>1e-5)[0]
print('bad tokens',len(bad))
for i in bad[:20]:
    t=toks[i]; print(i,t['name'],'rel %.1e'%rel[i],'S',len(t['invscale']),'nsub',list(t['nsub']),'aff',t['has_affine'],'off',t['off_page'],'nin',list(t['n_inbounds']),'sig %.2f'%t['sigma'])
# singles: each token alone (pair with empty second)
for i in bad[:6]:
    l1=run([toks[i]])[0]; print('alone',i,abs(l1-ref[i])/abs(ref[i]))
