import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, oracle_lib as O, opus_native_b200 as opn
lm,ch,pb=3,2,160
pk=opn.celt2_fill(500, 64, 2, 1, lm, ch, pb, transient_permille=250)[0]
payload=np.ascontiguousarray(pk[:,1:])
side,y,coef=opn.op_celt2_symbols(payload.reshape(-1), np.arange(64,dtype=np.uint32)*(pb-1), np.full(64,pb-1,np.uint32), lm, ch)
for s in range(3):
    w,_,wy,wc=O.celt2_decode_symbols(payload[s],lm,ch)
    d=side[s]["offsets"]
    print("dev ebits", list(d&255)); print("orc ebits", list(w.ebits))
    print("dev prio ", list((d>>8)&255)); print("orc prio ", list(w.fine_priority))
    print("dev pulses", list(d>>16)); print("orc pulses", list(w.pulses))
    print("dev bits_left", side[s]["balance"], "orc", 8*(pb-1)-(w.tell_frac+7)//8, "tell_frac", side[s]["tell_frac"], w.tell_frac)
    print("dev ff", side[s]["fine_final"].tolist()); print("orc ff", np.ctypeslib.as_array(w.fine_final).tolist())
