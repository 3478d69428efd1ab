"""CPU-only tests of the host-side mirror (pylabfea_b200.Model / Material): mesh numbering, boundary-condition
vectors and parameter setters, checked against fixtures written by the unmodified reference.  No device call."""
import numpy as np
import pytest

import pylabfea_b200 as FE
from test_gpu_model import cantilever, laminate, inclusion, uniax


def dense_K(g):
    n = int(g['Ndof'])
    K = np.zeros((n, n))
    K[g['K_row'], g['K_col']] = g['K_val']
    return K


@pytest.mark.parametrize('case,builder', [('cantilever10', lambda: cantilever(10)), ('laminate16x4', lambda: laminate()[0]),
                                          ('inclusion18', lambda: inclusion(18))])
def test_mesh_matches_reference(golden, case, builder):
    g = golden('kmat_' + case)
    fe = builder()
    assert fe.Ndof == int(g['Ndof']) and fe.Nel == len(g['nodes'])
    assert np.array_equal(fe.npos, g['npos'])
    assert np.array_equal(fe._nodes, g['nodes'])
    # the fixture numbers materials by first occurrence of the object, this model by section: compare the materials
    assert np.array_equal(np.array([fe.mat[i].E for i in fe._elem_mat]), g['mat_E'][g['mat_id']])
    assert np.array_equal(np.repeat(fe._dx, fe._NY), g['lx']) and np.all(g['ly'] == fe._dy)
    assert fe.element[3].nodes == list(g['nodes'][3]) and fe.element[3].Vel == g['lx'][3] * g['ly'][3] * float(g['thick'])


def test_uneven_sections_mesh(golden):
    g = golden('kmat_uneven4x5_plastic')
    fe = FE.Model(dim=2, planestress=False)
    fe.geom([2.5, 1.5], LY=3.)
    m1, m2 = FE.Material(), FE.Material()
    m1.elasticity(E=200.e3, nu=0.3)
    m2.elasticity(E=200.e3, nu=0.3)
    fe.assign([m1, m2])
    fe.mesh(NX=4, NY=5)
    assert np.array_equal(fe.npos, g['npos'])
    assert np.array_equal(np.repeat(fe._dx, 5), g['lx'])
    assert np.array_equal(fe._elem_mat, g['mat_id'])


@pytest.mark.parametrize('case,builder', [('cantilever10', lambda: cantilever(10)), ('laminate16x4', lambda: laminate()[0]),
                                          ('inclusion18_elastic', lambda: inclusion(18))])
def test_boundary_condition_vectors_reproduce_reference_solution(golden, case, builder):
    """_calc_BC (model.py:1070-1206 restated) + the reference's own dense K -> the reference's u, on the CPU."""
    gs = golden('solve_' + case)
    gk = golden('kmat_' + case.replace('_elastic', ''))
    fe = builder()
    K = dense_K(gk)
    dbcn = fe.bcn if fe.noset is not None else None
    fixed, du, w, fext = fe._calc_BC(fe.bcl, fe.bcb, fe.bcr, fe.bct, dbcn)
    free = np.nonzero(fixed == 0)[0]
    df = fext - K @ w
    du[free] = np.linalg.solve(K[np.ix_(free, free)], df[free])
    assert np.max(np.abs(du - gs['u'])) < 1e-12 * np.max(np.abs(gs['u']))
    assert np.max(np.abs(K @ du - gs['f'])) < 1e-9 * np.max(np.abs(gs['f'])) + 1e-9


def test_plasticity_setter_matches_reference(golden):
    g = golden('response_hillrv')
    m = FE.Material()
    m.elasticity(E=200.e3, nu=0.3)
    m.plasticity(sy=50., rv=[1.2, 1.0, 0.8, 1.0, 1.0, 1.0], sdim=6)
    assert np.array_equal(m.hill, g['hill']) and m.hill_6p and not m.hill_3p
    assert np.array_equal(m.CV, g['CV']) and m.khard == 0. and m.drucker == 0.
    g = golden('response_cubic')
    m = FE.Material()
    m.elasticity(C11=170.e3, C12=124.e3, C44=75.e3)
    assert m.E == float(g['E']) and m.nu == float(g['nu']) and np.array_equal(m.CV, g['CV'])
    with pytest.raises(ValueError):
        m.plasticity(sy=-1.)
    with pytest.raises(ValueError):
        m.elasticity(E=1.)
    with pytest.raises(ValueError):
        m.plasticity(sy=1., sdim=4)
    m.plasticity(sy=100., hill=[0.7, 1., 1.4], sdim=3)
    assert m.hill_3p and not m.hill_6p
    from pylabfea_b200 import _lib
    d = m._descriptor()     # sdim = 3: the principal-stress formulation has its own material kind
    assert d.kind == _lib.MAT_HILL3 and list(d.hill) == [0.7, 1., 1.4, 1., 1., 1.]


def test_stiffness_csr_wrapper():
    rowptr = np.array([0, 2, 3], dtype=np.int32)
    col = np.array([0, 1, 1], dtype=np.int32)
    val = np.array([2., -1., 3.])
    K = FE.model.StiffnessCSR(rowptr, col, val, 2)
    assert np.array_equal(np.asarray(K), [[2., -1.], [0., 3.]])
    assert np.array_equal(K @ np.array([1., 2.]), [0., 6.])
    assert K.tocsr().nnz == 3


def test_bc_argument_errors():
    fe = FE.Model(dim=2)
    with pytest.raises(TypeError):
        fe.bctop(0., 'spring')
    with pytest.raises(ValueError):
        fe.bcbot(1., 'force')
    with pytest.raises(ValueError):
        fe.bcright(1., 'disp', 'z')
    with pytest.raises(ValueError):
        FE.Model(dim=3)


def test_from_MLparam_parses_the_reference_export(golden):
    """Loader for the export_MLparam wire format (material.py:2133-2273) against a file written by the reference
    (tests/golden/mlparam_hill-svm.csv / -svm_meta.json, make_golden.py gen_mlparam): host-side parsing only."""
    import os
    import numpy as np
    import pylabfea_b200 as FE
    g = golden('mlparam_hill')
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    m = FE.Material.from_MLparam('mlparam_hill', path=here + os.sep)
    sv, dc, rho = m._svc
    assert np.array_equal(sv, g['sv']) and np.array_equal(dc, g['dual'])
    assert rho == float(g['rho']) and m.gam_yf == float(g['gamma']) and m.scale_seq == float(g['scale_seq'])
    assert m.sy == float(g['sy']) and m.dev_only == bool(g['dev_only']) and m.ML_yf
    assert np.allclose(m.CV, g['CV'], rtol=1e-15, atol=0.)
    assert m.meta['Data']['Class'] == 'SVC_parameters' and m.meta['Model']['Creator'] == 'pylabfea'


def test_sdim3_svc_descriptor(golden):
    """An SVC trained in the sdim = 3 feature space crosses the C-ABI as PLF_MAT_SVC3 with a 6-column table."""
    from pylabfea_b200 import _lib
    g = golden('svc3_mlplast')
    m = FE.Material()
    m.elasticity(E=200.e3, nu=0.3)
    m.plasticity(sy=150., sdim=3)
    with pytest.raises(ValueError):
        m.set_svc(np.zeros((4, 6)), np.zeros(4), 0., 1.)        # sdim = 3 models have 2 features
    m.set_svc(g['sv'], g['dual'], float(g['rho']), float(g['gamma']))
    d = m._descriptor()
    assert d.kind == _lib.MAT_SVC3 and d.nsv == len(g['dual']) and d.svc_scale_seq == 150. and m.Ndof == 2
    sv = np.ctypeslib.as_array(d.svc_sv, shape=(d.nsv, 6))
    assert np.array_equal(sv[:, 0:2], g['sv']) and not sv[:, 2:].any()
    m6 = FE.Material()
    m6.elasticity(E=200.e3, nu=0.3)
    m6.plasticity(sy=150., sdim=6)
    with pytest.raises(ValueError):
        m6.set_svc(g['sv'], g['dual'], float(g['rho']), float(g['gamma']))


def test_from_reference_duck_typed_sdim3_ml_material(golden):
    """Material.from_reference copies what a pyLabFEA Material carries (here a stand-in object with the attributes the
    reference sets in plasticity() and setup_yf_SVM_3D()) -- sdim = 3 with a trained SVC included."""
    from types import SimpleNamespace
    from pylabfea_b200 import _lib
    g = golden('svc3_mlplast')
    svc = SimpleNamespace(support_vectors_=g['sv'], dual_coef_=g['dual'][None, :], intercept_=np.array([float(g['rho'])]),
                          _gamma=float(g['gamma']), gamma=float(g['gamma']))
    ref = SimpleNamespace(name='ML flow rule', num=1, E=200.e3, nu=0.3, C11=g['CV'][0, 0], C12=g['CV'][0, 1], C44=g['CV'][3, 3],
                          CV=g['CV'], sy=150., sy0=150., khard=0., sdim=3, drucker=0., lhs=None, hill=np.ones(3), hill_6p=False,
                          hill_3p=False, tresca=False, barlat=False, ML_yf=True, whdat=False, txdat=False, svm_yf=svc,
                          scale_seq=float(g['scale_seq']), dev_only=False)
    m = FE.Material.from_reference(ref)
    d = m._descriptor()
    assert d.kind == _lib.MAT_SVC3 and d.nsv == len(g['dual']) and m.sdim == 3 and m.Ndof == 2
    assert d.svc_gamma == float(g['gamma']) and d.svc_rho == float(g['rho']) and np.array_equal(m.CV, g['CV'])
    ref.sdim = 4
    with pytest.raises(ValueError):
        FE.Material.from_reference(ref)


def test_from_MLparam_parses_an_sdim3_export(golden):
    """export_MLparam of an sdim = 3 model (Ndof = 2, written by the reference: tests/golden/mlparam_svc3-svm.csv,
    make_golden.py gen_mlparam3) loads as an sdim = 3 material with the same SVC."""
    import os
    from pylabfea_b200 import _lib
    g = golden('mlparam_svc3')
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    m = FE.Material.from_MLparam('mlparam_svc3', path=here + os.sep)
    sv, dc, rho = m._svc
    assert m.sdim == 3 and m.Ndof == 2 and sv.shape == (len(g['dual']), 6)
    assert np.array_equal(sv[:, 0:2], g['sv']) and not sv[:, 2:].any() and np.array_equal(dc, g['dual'])
    assert rho == float(g['rho']) and m.gam_yf == float(g['gamma']) and m.scale_seq == float(g['scale_seq']) == m.sy
    assert np.allclose(m.CV, g['CV'], rtol=1e-15, atol=0.)
    assert m._descriptor().kind == _lib.MAT_SVC3
