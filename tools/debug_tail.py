import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import ft2_density
from rayrtrace_b200 import ft2
from rayrtrace_b200.beam import Beam, field_result, field_r_device, receiver_field_device
from rayrtrace_b200.nfield import Nfield
fx = np.load(os.path.join(ROOT, "tests/golden/ft2_phases.npz"))
n, x, y = ft2_density(int(fx["nx"]), float(fx["n0"]))
nf = Nfield(n * 4.5, x=x, y=y)
Beam._nray = int(fx["nray"])
ant_tx = ft2.ants() * np.array([1., -1.]); ant_rx = ant_tx * np.array([1., -1.])
tx_, rx_ = ant_tx[4], ant_rx[4]
rx = Beam(x0=rx_[0], y0=rx_[1], kx0=0, ky0=-1)
tx = Beam(x0=tx_[0], y0=tx_[1], kx0=0, ky0=1)
tx.rays(n_field=nf, keep_traj=False, receiver=(float(rx_[1]), "y"))
h = field_result(tx, rx, device=True)
g = field_result(tx, rx, device="gpu")
print("host x_max", h[2], "ph_max", h[3], "props", h[5], "signal", h[4])
print("gpu  x_max", g[2], "ph_max", g[3], "props", g[5], "signal", g[4])
print("n", len(h[1]), len(g[1]), "x equal", np.array_equal(h[1], g[1]), "max |field diff|", np.abs(h[0] - g[0]).max(), np.abs(h[0]).max())
print("x range", h[1][0], h[1][-1])
print("want", fx["dense_ph_by_max"][1, 4])
import ctypes, torch
from rayrtrace_b200 import _cabi
from rayrtrace_b200.lmfit import lmdif3, fd_steps
r = receiver_field_device(tx, float(rx_[1]))
xs = torch.from_numpy(g[1]).cuda(); ys = torch.from_numpy(np.abs(g[0])).cuda()
out = torch.empty(10, dtype=torch.float64, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def sums(p, h):
    hs = (ctypes.c_double * 3)(*[float(v) for v in h])
    _cabi.check(_cabi.lib.stb_gauss_fit_sums(xs.data_ptr(), ys.data_ptr(), len(g[1]), float(p[0]), float(p[1]), float(p[2]), hs, out.data_ptr(), st))
    s = out.cpu().numpy()
    return np.array([[s[0], s[1], s[2]], [s[1], s[3], s[4]], [s[2], s[4], s[5]]]), s[6:9].copy(), float(s[9])
def model(xa,amp,xc,s): return amp*np.exp(-(xa-xc)**2/2/s**2)
def sums_np(p,h):
    x=g[1]; y=np.abs(g[0]); gg=model(x,*p); rr=gg-y
    J=np.stack([(model(x,p[0]+h[0],p[1],p[2])-gg)/h[0],(model(x,p[0],p[1]+h[1],p[2])-gg)/h[1],(model(x,p[0],p[1],p[2]+h[2])-gg)/h[2]],1)
    return J.T@J, J.T@rr, float(rr@rr)
p0 = (np.abs(g[0]).max(), g[2], tx.waist)
print("p0", p0)
print("dev sums", sums(np.array(p0), fd_steps(p0)))
print("np  sums", sums_np(np.array(p0), fd_steps(p0)))
print("lmdif dev", lmdif3(sums, p0))
print("lmdif np ", lmdif3(sums_np, p0))
