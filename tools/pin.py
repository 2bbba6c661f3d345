import sys, os, numpy as np
sys.path.insert(0, '/root/repo/oracle')
import ref_run, oracle

def pin(cfg, pgen, inp, ov, solver, mhd, ncyc=4, quiet=False):
    ov = dict(ov); ov['time/nlim'] = ncyc
    res = ref_run.run_reference(cfg, pgen, inp, ov, rst_every_cycle=True)
    rsts = [ref_run.read_rst(p) for p in res['rst']]
    par = rsts[0]['par']
    P = oracle.params_from_athinput(par, mhd, solver, ng=rsts[0]['nghost'])
    M = oracle.OracleMesh(P)
    M.load_rst(rsts[0])
    M.initialize()
    ok = True
    print(cfg, pgen, 'nb', M.nb, 'dt0 ref %.17e oracle %.17e' % (rsts[0]['dt'], M.dt), 'EQ' if rsts[0]['dt']==M.dt else 'DIFF')
    ok &= rsts[0]['dt']==M.dt
    for c in range(1, len(rsts)):
        M.cycle()
        r = rsts[c]
        mx = 0.0
        for blk in r['blocks']:
            b = M.block_of(*blk['loc'][:3])
            i = M.info[b]
            sl = (slice(None), slice(i['ks'], i['ke']+1), slice(i['js'], i['je']+1), slice(i['is'], i['ie']+1))
            d = np.abs(M.array(b,'u')[sl] - blk['u'][sl]).max()
            mx = max(mx, d)
            full = np.array_equal(M.array(b,'u'), blk['u'])
            if mhd:
                for nm in ('b1','b2','b3'):
                    d = np.abs(M.array(b,nm) - blk[nm]).max(); mx = max(mx, d)
        eq = (r['dt'] == M.dt)
        ok &= eq and mx == 0.0
        print('  cycle', c, 'dt ref %.17e or %.17e' % (r['dt'], M.dt), 'EQ' if eq else 'DIFF', 'max|du|', mx, 'ghosts-equal', full)
    ref_run.cleanup(res)
    return ok

if __name__ == '__main__' and len(sys.argv)==1:
    small = {'mesh/nx1':16,'mesh/nx2':8,'mesh/nx3':8,'meshblock/nx1':16,'meshblock/nx2':8,'meshblock/nx3':8}
    pin('mhd_hlld_ng2','linear_wave','/root/repo/inputs/athinput.linear_wave3d', small, 'hlld', True)

def all_pins():
    I='/root/repo/inputs/'
    res = {}
    lw = {'mesh/nx1':16,'mesh/nx2':8,'mesh/nx3':8}
    res['lw 8blocks'] = pin('mhd_hlld_ng2','linear_wave',I+'athinput.linear_wave3d', dict(lw, **{'meshblock/nx1':8,'meshblock/nx2':4,'meshblock/nx3':4}), 'hlld', True)
    bl = {'mesh/nx1':16,'mesh/nx2':16,'mesh/nx3':16, 'problem/radius':0.3}
    res['blast 1 block'] = pin('mhd_hlld_ng2','blast',I+'athinput.blast', dict(bl, **{'meshblock/nx1':16,'meshblock/nx2':16,'meshblock/nx3':16}), 'hlld', True, ncyc=6)
    res['blast 8 blocks'] = pin('mhd_hlld_ng2','blast',I+'athinput.blast', dict(bl, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8}), 'hlld', True, ncyc=6)
    ot = {'mesh/nx1':32,'mesh/nx2':32}
    res['OT ppm 4 blocks'] = pin('mhd_hlld_ng3','orszag_tang',I+'athinput.orszag_tang', dict(ot, **{'meshblock/nx1':16,'meshblock/nx2':16}), 'hlld', True, ncyc=5)
    res['OT plm 1 block'] = pin('mhd_hlld_ng2','orszag_tang',I+'athinput.orszag_tang', dict(ot, **{'meshblock/nx1':32,'meshblock/nx2':32,'time/xorder':2}), 'hlld', True, ncyc=5)
    kh = {'mesh/nx1':16,'mesh/nx2':16,'mesh/nx3':16}
    res['KH rk2 ppm 8 blocks'] = pin('hydro_hllc_ng3','kh',I+'athinput.kh', dict(kh, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8}), 'hllc', False, ncyc=5)
    res['sod'] = pin('hydro_hllc_ng2','shock_tube',I+'athinput.sod', {'mesh/nx1':64,'meshblock/nx1':32}, 'hllc', False, ncyc=8)
    res['sod hlle'] = pin('hydro_hlle_ng2','shock_tube',I+'athinput.sod', {'mesh/nx1':64,'meshblock/nx1':64}, 'hlle', False, ncyc=8)
    res['sod roe'] = pin('hydro_roe_ng2','shock_tube',I+'athinput.sod', {'mesh/nx1':64,'meshblock/nx1':64}, 'roe', False, ncyc=8)
    res['lw mhd hlle'] = pin('mhd_hlle_ng2','linear_wave',I+'athinput.linear_wave3d', dict(lw, **{'meshblock/nx1':16,'meshblock/nx2':8,'meshblock/nx3':8,'problem/amp':0.1}), 'hlle', True)
    res['lw mhd roe'] = pin('mhd_roe_ng2','linear_wave',I+'athinput.linear_wave3d', dict(lw, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8,'problem/amp':0.1}), 'roe', True)
    print(res)
if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1]=='all':
    all_pins()

def lhll_pins():
    I='/root/repo/inputs/'
    res={}
    bl = {'mesh/nx1':16,'mesh/nx2':16,'mesh/nx3':16, 'problem/radius':0.3}
    res['lhlld blast 8blk'] = pin('mhd_lhlld_ng2','blast',I+'athinput.blast', dict(bl, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8}), 'lhlld', True, ncyc=6)
    ot = {'mesh/nx1':32,'mesh/nx2':32,'time/xorder':2}
    res['lhlld OT 4blk'] = pin('mhd_lhlld_ng2','orszag_tang',I+'athinput.orszag_tang', dict(ot, **{'meshblock/nx1':16,'meshblock/nx2':16}), 'lhlld', True, ncyc=5)
    res['lhllc blast 8blk'] = pin('hydro_lhllc_ng2','blast',I+'athinput.blast', dict(bl, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8}), 'lhllc', False, ncyc=6)
    res['lhllc sod'] = pin('hydro_lhllc_ng2','shock_tube',I+'athinput.sod', {'mesh/nx1':64,'meshblock/nx1':32}, 'lhllc', False, ncyc=8)
    kh = {'mesh/nx1':16,'mesh/nx2':16,'mesh/nx3':1,'time/xorder':2,'time/integrator':'vl2'}
    res['lhllc kh 2d'] = pin('hydro_lhllc_ng2','kh',I+'athinput.kh', dict(kh, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':1}), 'lhllc', False, ncyc=5)
    print(res)
if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1]=='lhll':
    lhll_pins()

def refl_pins():
    I='/root/repo/inputs/'
    res={}
    R={'mesh/ix1_bc':'reflecting','mesh/ox1_bc':'reflecting','mesh/ix2_bc':'reflecting','mesh/ox2_bc':'reflecting','mesh/ix3_bc':'reflecting','mesh/ox3_bc':'reflecting'}
    bl = {'mesh/nx1':16,'mesh/nx2':16,'mesh/nx3':16, 'problem/radius':0.6}
    res['refl blast mhd 8blk'] = pin('mhd_hlld_ng2','blast',I+'athinput.blast', dict(bl, **R, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8}), 'hlld', True, ncyc=8)
    R2={'mesh/ix1_bc':'reflecting','mesh/ox1_bc':'outflow','mesh/ix2_bc':'periodic','mesh/ox2_bc':'periodic','mesh/ix3_bc':'outflow','mesh/ox3_bc':'reflecting'}
    res['mixed blast hydro ppm 8blk'] = pin('hydro_hllc_ng3','blast',I+'athinput.blast', dict(bl, **R2, **{'meshblock/nx1':8,'meshblock/nx2':8,'meshblock/nx3':8,'time/xorder':3}), 'hllc', False, ncyc=8) if os.path.exists('/root/repo/oracle/_ref/hydro_hllc_ng3/athena_blast') else None
    res['refl sod'] = pin('hydro_hllc_ng2','shock_tube',I+'athinput.sod', {'mesh/nx1':64,'meshblock/nx1':32,'mesh/ix1_bc':'reflecting','mesh/ox1_bc':'reflecting','time/tlim':1.0}, 'hllc', False, ncyc=12)
    print(res)
if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1]=='refl':
    refl_pins()
