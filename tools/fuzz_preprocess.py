"""Fuzz of the pre-processing oracle and of the emulated device code (tests/native/preprocess_emu.cpp) against Pillow
itself: random input sizes (a fifth of them 1-4 pixels wide), random targets, random crop windows, random / saturated
pixels, for --seconds of wall clock.  CPU only.  This is how the height-first rule of Pillow >= 11 was found.

    g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC -o /tmp/libpreemu.so tests/native/preprocess_emu.cpp
    python tools/fuzz_preprocess.py [--seconds 150]
"""
import argparse
import os
import ctypes as C, numpy as np, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image
from oracle import preprocess_oracle as po
lib=C.CDLL('/tmp/libpreemu.so')
rng=np.random.default_rng(12345)
ap = argparse.ArgumentParser(); ap.add_argument('--seconds', type=float, default=150.0); args = ap.parse_args()
t0=time.time(); n=0; bad=0
while time.time()-t0 < args.seconds:
    h,w=int(rng.integers(1,400)),int(rng.integers(1,400)) if rng.random()<0.8 else int(rng.integers(1,5))
    oh,ow=int(rng.integers(1,300)),int(rng.integers(1,300))
    img=rng.integers(0,256,(h,w,3),dtype=np.uint8)
    if rng.random()<0.3: img=(img>127).astype(np.uint8)*255
    ref=np.asarray(Image.fromarray(img).resize((ow,oh),Image.BICUBIC))
    # random crop window inside the resized image
    ch,cw=int(rng.integers(1,oh+1)),int(rng.integers(1,ow+1)); top,left=int(rng.integers(0,oh-ch+1)),int(rng.integers(0,ow-cw+1))
    out=np.full((1,3,ch,cw),np.nan,np.float32)
    rc=lib.emu_image_preprocess(np.ascontiguousarray(img[None]).ctypes.data_as(C.c_void_p),1,h,w,oh,ow,top,left,ch,cw,(C.c_float*3)(.5,.5,.5),(C.c_float*3)(.5,.5,.5),out.ctypes.data_as(C.c_void_p),None)
    orc=po.resize_bicubic_u8(img,oh,ow)
    if not np.array_equal(orc,ref): bad+=1; print("ORACLE MISMATCH",h,w,oh,ow)
    if po.vertical_pass_first(h,w,oh):
        assert rc==-5,(rc,h,w,oh,ow); n+=1; refused=globals().get('refused',0)+1; globals()['refused']=refused; continue
    assert rc==0,(rc,h,w,oh,ow)
    want=po.to_tensor_normalize(ref[top:top+ch,left:left+cw],(.5,.5,.5),(.5,.5,.5))
    if not np.array_equal(out[0],want):
        bad+=1; print("MISMATCH",h,w,oh,ow,top,left,ch,cw)
    n+=1
print("cases",n,"bad",bad,"refused",globals().get("refused",0))
