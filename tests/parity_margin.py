import sys, importlib, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
pkg = importlib.import_module("navierstokes-capoferri_cecchettini_untila_b200")
from oracle import ns_oracle
import test_gpu_parity as T
for rep in range(4):
    prob, orc, dim, nu, um = conftest.make_case(pkg, ns_oracle, "3d-cylinder")
    dev = T._device(pkg, prob, dim, nu)
    orc.set_solver(1e-12, 30, 10000, 1e-10)
    dev.set_solver(gmres_rtol=1e-12, restart=60)
    t=0.0
    for step in range(3):
        t+=0.01
        orc.assemble(t); dev.assemble(t)
        rc,it_o,_,_=orc.solve_time_step(); it_d,_,_=dev.solve_time_step()
        xo,xd=orc.solution(),dev.solution()
        f_o=orc.compute_forces(t); f_d=dev.compute_forces(prob.mean_velocity(t))
        print(rep,step,"its",it_o,it_d,"err",np.linalg.norm(xd-xo)/np.linalg.norm(xo),"dcd",abs(f_d[2]-f_o[2]),"dcl",abs(f_d[3]-f_o[3]),flush=True)
        dev.set_solution(xo)
