"""Host graph layer vs. golden vectors exported from the reference
(tests/golden/graphs.npz, made by oracle/gen_golden.py)."""
import random

import numpy as np
import pytest

from conftest import load_golden, make_code, golden_names

META, ARR = load_golden('graphs')


@pytest.mark.parametrize('k', range(len(META['cases'])), ids=lambda k: str(META['cases'][k]['spec']))
def test_flat_graph_equals_reference(k):
    case = META['cases'][k]
    code = make_code(case['spec'])
    g = code.FLAT
    assert g.n_nodes == case['N'] and g.n_boundary == case['B'] and g.n_edges == case['E']
    assert code.N_EDGES == case['n_edges_attr']
    assert code.DIMENSION == case['dimension']
    assert code.SCHEME.WINDOW_HEIGHT == case['height']
    np.testing.assert_array_equal(g.edge_u, ARR[f'g{k}_edge_u'])
    np.testing.assert_array_equal(g.edge_v, ARR[f'g{k}_edge_v'])
    np.testing.assert_array_equal(g.edge_class, ARR[f'g{k}_edge_class'])
    np.testing.assert_array_equal(np.array(code.NODES, np.int32), ARR[f'g{k}_coords'])
    for p, want in case['probs'].items():
        got = code.NOISE.class_probabilities(float(p))
        assert got == pytest.approx(want, rel=1e-14, abs=0)
    if f'g{k}_commit_edges' in ARR:
        got = sorted(g.edge_index[e] for e in code.SCHEME._COMMIT_EDGES)
        np.testing.assert_array_equal(got, ARR[f'g{k}_commit_edges'])


def test_boundary_ids_lowest_and_csr_sorted():
    code = make_code(dict(code='Surface', d=5, noise='circuit-level'))
    g = code.FLAT
    assert all(code.is_boundary(v) for v in code.NODES[:g.n_boundary])
    assert not any(code.is_boundary(v) for v in code.NODES[g.n_boundary:])
    for v in range(g.n_nodes):
        sl = g.inc_edge[g.inc_off[v]:g.inc_off[v + 1]]
        assert list(sl) == sorted(sl)
        assert len(sl) == len(code.INCIDENT_EDGES[code.NODES[v]])
    assert g.max_degree == 12


def test_neighbour_table_matches_edges():
    for spec in (dict(code='Surface', d=3, noise='code capacity'),
                 dict(code='Surface', d=3, noise='phenomenological'),
                 dict(code='Surface', d=3, noise='circuit-level')):
        code = make_code(spec)
        g = code.FLAT
        seen = set()
        for v in range(g.n_nodes):
            for s in range(g.nbr_edge.shape[1]):
                e = g.nbr_edge[v, s]
                if e < 0:
                    continue
                w = g.nbr_node[v, s]
                if g.nbr_owned[v, s]:
                    assert (g.edge_u[e], g.edge_v[e]) == (v, w)
                else:
                    assert (g.edge_u[e], g.edge_v[e]) == (w, v)
                seen.add((v, e))
        assert len(seen) == 2 * g.n_edges      # every half-edge has exactly one direction slot


@pytest.mark.parametrize('name', [n for n in golden_names('uf_') if 'fwd' not in n])
def test_host_make_error_matches_reference_stream(name):
    """random.seed(s) + Code.make_error reproduces the reference's error sets
    (noise/main.py:114-115,309-314) and syndromes (_base_classes.py:427-450)."""
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    g = code.FLAT
    random.seed(meta['seed'])
    for s in range(min(meta['shots'], 50)):
        error = code.make_error(meta['p'])
        np.testing.assert_array_equal(g.edges_to_bits(error), arr['err_bits'][s])
        np.testing.assert_array_equal(g.syndrome_to_bits(code.get_syndrome(error)), arr['syn_bits'][s])
        assert g.bits_to_edges(arr['err_bits'][s]) == error
        assert code.get_logical_error(error ^ g.bits_to_edges(arr['default_corr'][s])) == arr['default_logical'][s]


def test_constructor_errors():
    import localuf_b200 as L
    with pytest.raises(ValueError):
        L.Surface(3, 'code capacity', window_height=3)
    with pytest.raises(ValueError):
        L.Surface(3, 'phenomenological', commit_height=1)
    with pytest.raises(TypeError):
        L.Surface(3, 'code capacity', scheme='forward')
    with pytest.raises(ValueError):
        L.Surface(3, 'phenomenological', scheme='frugal', window_height=4)
    with pytest.raises(NotImplementedError):
        L.Surface(3, 'phenomenological', scheme='forward', merge_equivalent_boundary_nodes=True)
    with pytest.raises(NotImplementedError):
        L.Repetition(3, 'circuit-level')


def test_known_counts():
    """Edge/node counts quoted in SURVEY.md section 8 for the BASELINE configs."""
    import localuf_b200 as L
    c = L.Repetition(5, 'code capacity')
    assert (c.NODE_COUNT, c.DETECTOR_COUNT, c.N_EDGES) == (6, 4, 5)
    c = L.Surface(9, 'code capacity')
    assert (c.NODE_COUNT, c.DETECTOR_COUNT, c.N_EDGES) == (90, 72, 145)
    c = L.Surface(11, 'phenomenological')
    assert (c.NODE_COUNT, c.DETECTOR_COUNT, c.N_EDGES) == (1452, 1210, 3531)
    assert c.FLAT.max_degree == 6
