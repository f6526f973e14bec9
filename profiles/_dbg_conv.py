import ctypes as C, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from loco_edit_b200 import _lib
from loco_edit_b200._lib import check, ptr, stream_ptr
lib=_lib.load(); dev=torch.device("cuda:0")
for (kind,N,H,W,Cin,Cout) in [(0,6,256,256,128,128),(0,6,256,256,256,128),(0,6,64,64,256,256)]:
    ksz = 1 if kind==1 else 3
    x=torch.randn(N,H,W,Cin,device=dev); wp=torch.randn(Cout*Cin*ksz*ksz,device=dev)*0.01; y=torch.empty(N,H,W,Cout,device=dev)
    scr=torch.zeros(32<<20,dtype=torch.uint8,device=dev)
    ms,ks,gr=C.c_float(),C.c_int(),C.c_int()
    check(lib.loco_conv_bench(kind,ptr(x),N,H,W,Cin,ptr(wp),Cout,Cin,ptr(y),ptr(scr),scr.numel(),1,30,C.byref(ms),C.byref(ks),C.byref(gr),stream_ptr()),"b")
    fl=2.0*N*H*W*Cout*Cin*ksz*ksz
    print("debug",os.environ.get("LOCO_CONV_DEBUG","0"),"NT",os.environ.get("LOCO_CONV_NT","auto"), (kind,N,H,W,Cin,Cout), f"{ms.value*1e3:.1f} us  {fl/ms.value/1e9:.0f} TF/s")
