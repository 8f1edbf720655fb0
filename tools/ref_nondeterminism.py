"""Development aid: shows that the REFERENCE ITSELF (unmodified GLSL on llvmpipe) is not run-to-run deterministic on pixels whose
nearest hit is an analytic light: pathtrace.glsl:246-253 evaluates GetMaterialsAndTextures on a State whose matID / triID were
never written, and llvmpipe does not initialise shader temporaries.  Measured on c3mini, 4 spp, two runs of the same binary
(default threads vs LP_NUM_THREADS=3): 871 pixels differ, 862 of them emitter-first-hit pixels (of 9136).  Those pixels are
therefore excluded from the bit-identity checks of multi-sample images in tests/test_oracle_golden.py.
"""
import os, sys, subprocess, tempfile, json, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT)
from scenes import gen_scenes
REFBIN=ROOT+'/oracle/_ref/lf_ref_llvmpipe'
with tempfile.TemporaryDirectory() as tmp:
    scene=gen_scenes.c3_mini(os.path.join(tmp,'assets'))
    imgs=[]
    for k in range(2):
        out=os.path.join(tmp,f'o{k}.f32')
        env=gen_scenes.llvmpipe_env(threads=(None if k==0 else 3))
        r=subprocess.run([REFBIN,'--scene',scene,'--spp','4','--out',out,'--timing-json'],env=env,check=True,capture_output=True,text=True)
        info=json.loads(r.stdout.strip().splitlines()[-1])
        imgs.append(np.fromfile(out,np.float32).reshape(info['height'],info['width'],3))
    g=np.load(ROOT+'/tests/golden/c3mini_llvmpipe.npz')
    em=g['hits_emitter']>0
    d=(imgs[0]!=imgs[1]).any(axis=2)
    print('pixels differing between two runs of the same binary:', d.sum(), 'of which first hit is an emitter:', (d&em).sum(), 'emitter pixels', em.sum())
