import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import torch
from loco_edit_b200 import _lib
from loco_edit_b200._lib import check, ptr, stream_ptr
lib=_lib.load(); dev=torch.device("cuda:0")
for (kind,N,H,W,Cin,Cout) in [(1,1,8,16,32,128),(0,1,16,16,512,512)]:
    ksz = 1 if kind==1 else 3
    x=torch.randn(N,H,W,Cin,device=dev); wp=torch.randn(Cout*Cin*ksz*ksz,device=dev)*0.01; y=torch.empty(N,H,W,Cout,device=dev)
    scr=torch.zeros(32<<20,dtype=torch.uint8,device=dev)
    ms,ks,gr=C.c_float(),C.c_int(),C.c_int()
    check(lib.loco_conv_bench(kind,ptr(x),N,H,W,Cin,ptr(wp),Cout,Cin,ptr(y),ptr(scr),scr.numel(),1,200,C.byref(ms),C.byref(ks),C.byref(gr),stream_ptr()),"b")
    print(os.environ.get("LOCO_CONV_DEBUG","0"), (kind,N,H,W,Cin,Cout), "grid",gr.value, f"{ms.value*1e3:.2f} us")
