"""Iteration counts of candidate solver configurations against grid size (CPU oracle): why the bench uses MG + mass Schur."""
import sys, time; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(R,'oracle'))
import sp_oracle as so
base="-ksp_rtol 1e-8 -ksp_max_it 3000 -pc_type fieldsplit -pc_fieldsplit_type schur "
cfgs={
 "fgmres_upper_massQ_mg": "-ksp_type fgmres -pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels %d -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi ",
 "fgmres_upper_lsc_mg(cheb8 on L)": "-ksp_type fgmres -pc_fieldsplit_schur_fact_type upper -pc_fieldsplit_schur_precondition self -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type mg -fieldsplit_0_pc_mg_levels %d -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type lsc -fieldsplit_1_pc_lsc_scale_diag -fieldsplit_1_lsc_ksp_type chebyshev -fieldsplit_1_lsc_ksp_max_it 8 -fieldsplit_1_lsc_pc_type jacobi ",
 "gmres_full_massQ_jacobiA00": "-ksp_type gmres -pc_fieldsplit_schur_fact_type full -pc_fieldsplit_schur_precondition user -fieldsplit_0_ksp_type preonly -fieldsplit_0_pc_type jacobi -fieldsplit_1_ksp_type preonly -fieldsplit_1_pc_type jacobi ",
}
for nx in (16,32,64,128):
    lev={16:2,32:3,64:4,128:5}[nx]
    p=so.Problem(nx,nx,kkt=True,rhs_kind=1)
    for name,o in cfgs.items():
        opts=base+(o % lev if '%d' in o else o)
        t=time.time(); r=so.Solver(p,opts).solve(history=False); 
        print(nx, name, r['its'], r['reason'], round(time.time()-t,2), flush=True)
