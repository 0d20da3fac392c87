"""permute_subentity_closure / _inv (reference finite-element.h:860-1035, tables finite-element.cpp:1449-1705).

CPU part: the numpy restatement (oracle/restate.py) against the reference's own test vectors
(test/test_dof_transformations.py:480-600) and the compiled reference; the library's host tables against the
compiled reference.  GPU part: the batched kernel through the C ABI, bit-exact against the compiled reference.
"""

import numpy as np
import pytest

from conftest import make_pair

P = 1
POINT, INTERVAL, TRI, TET, QUAD, HEX, PRISM, PYRAMID = 0, 1, 2, 3, 4, 5, 6, 7
EQ, GLL = 1, 2

# (cell, degree, sub-entity type, {entity_info: result}) -- copied from the reference's test parametrisation
REFERENCE_VECTORS = [
    (TET, 3, TRI, {0: [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], 2: [1, 2, 0, 6, 5, 8, 7, 3, 4, 9], 4: [2, 0, 1, 7, 8, 4, 3, 6, 5, 9],
                   1: [0, 2, 1, 4, 3, 7, 8, 5, 6, 9], 3: [1, 0, 2, 5, 6, 3, 4, 8, 7, 9], 5: [2, 1, 0, 8, 7, 6, 5, 4, 3, 9]}),
    (TET, 4, TRI, {0: list(range(15)), 2: [1, 2, 0, 8, 7, 6, 11, 10, 9, 3, 4, 5, 13, 14, 12],
                   4: [2, 0, 1, 9, 10, 11, 5, 4, 3, 8, 7, 6, 14, 12, 13], 1: [0, 2, 1, 5, 4, 3, 9, 10, 11, 6, 7, 8, 12, 14, 13],
                   3: [1, 0, 2, 6, 7, 8, 3, 4, 5, 11, 10, 9, 13, 12, 14], 5: [2, 1, 0, 11, 10, 9, 8, 7, 6, 5, 4, 3, 14, 13, 12]}),
    (TET, 1, INTERVAL, {0: [0, 1], 1: [1, 0]}),
    (TET, 2, INTERVAL, {0: [0, 1, 2], 1: [1, 0, 2]}),
    (TET, 3, INTERVAL, {0: [0, 1, 2, 3], 1: [1, 0, 3, 2]}),
]

SUBS = {TRI: [INTERVAL], QUAD: [INTERVAL], TET: [INTERVAL, TRI], HEX: [INTERVAL, QUAD], PRISM: [INTERVAL, TRI, QUAD],
        PYRAMID: [INTERVAL, TRI, QUAD]}


def restated(r):
    from oracle import restate as R

    arrs = {"meta": np.array([r.cell_type, r.polyset_type, r.degree, r.embedded_superdegree, r.map_type]),
            "value_shape": np.array(r.value_shape, dtype=np.int64), "coefficient_matrix": r.coefficient_matrix,
            "dof_ordering": r.dof_ordering}
    flat, offs = [], [0]
    for row in r.entity_dofs:
        for e in row:
            flat += e
            offs.append(len(flat))
    arrs["entity_dof_offsets"], arrs["entity_dofs_flat"] = np.array(offs), np.array(flat if flat else [0])
    for ct, v in r.entity_transformations.items():
        arrs[f"etrans_{ct}"] = v
    return R.RestatedElement(arrs)


def closure_size(r, sub):
    from oracle.restate import CLOSURE_CONN

    conn = CLOSURE_CONN[(r.cell_type, sub)]
    return sum(len(r.entity_dofs[dim][i]) for dim, ents in enumerate(conn) for i in ents)


@pytest.mark.parametrize("cell,degree,sub,results", REFERENCE_VECTORS)
def test_restatement_reference_vectors(ref, cell, degree, sub, results):
    e = restated(ref.RefElement(P, cell, degree, EQ))
    for info, want in results.items():
        d = list(range(len(want)))
        assert list(e.permute_subentity_closure(d, info, sub)) == want
        assert list(e.permute_subentity_closure(d, info, sub, inverse=True)) == list(range(len(want)))


@pytest.mark.parametrize("cell", [TRI, QUAD, TET, HEX, PRISM, PYRAMID])
@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5])
def test_restatement_and_host_tables_vs_compiled_reference(ref, cell, degree):
    try:
        r, e = make_pair(ref, P, cell, degree, GLL if degree > 2 else 0)
    except RuntimeError:
        pytest.skip("element not available in the reference")
    o = restated(r)
    for sub in SUBS[cell]:
        m = closure_size(r, sub)
        assert e.subentity_closure_size(sub) == m
        for info in range(2 if sub == INTERVAL else 8):
            for inverse in (False, True):
                d = (np.arange(m, dtype=np.int32) + 100)[None].copy()
                want = r.permute_subentity_closure_batch(d.copy(), np.array([info], dtype=np.uint32), sub, None, inverse)[0]
                assert np.array_equal(np.array(o.permute_subentity_closure(list(d[0]), info, sub, inverse)), want)
                assert np.array_equal(d[0][e.subentity_closure_map(sub, info, inverse)], want), (sub, info, inverse)


def test_host_tables_with_dof_ordering_and_native_factory(ref):
    """A custom dof_ordering renumbers entity dofs but not the closure-local numbering; the native factory builds
    the same tables as an element described by the reference's arrays."""
    import basix_b200 as bx

    order = [int(v) for v in np.random.default_rng(7).permutation(20)]
    r, e = make_pair(ref, P, TET, 3, GLL, dof_ordering=order)
    native = bx.create_element(P, TET, 3, GLL, dof_ordering=order)
    for sub in (INTERVAL, TRI):
        m = closure_size(r, sub)
        for info in range(2 if sub == INTERVAL else 8):
            for inverse in (False, True):
                d = np.arange(m, dtype=np.int32)[None].copy()
                want = r.permute_subentity_closure_batch(d.copy(), np.array([info], dtype=np.uint32), sub, None, inverse)[0]
                assert np.array_equal(d[0][e.subentity_closure_map(sub, info, inverse)], want)
                assert np.array_equal(d[0][native.subentity_closure_map(sub, info, inverse)], want)


# ---- GPU: the batched kernel through the C ABI ------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cell,degree", [(TRI, 3), (QUAD, 4), (TET, 1), (TET, 2), (TET, 5), (HEX, 2), (HEX, 4), (PRISM, 3),
                                         (PYRAMID, 2)])
def test_closure_batch_matches_reference(ref, cell, degree):
    import torch

    r, e = make_pair(ref, P, cell, degree, GLL if degree > 2 else 0)
    rng = np.random.default_rng(degree)
    nents = {INTERVAL: len(r.entity_dofs[1]), TRI: len(r.entity_dofs[2]) if r.tdim == 3 else 0,
             QUAD: len(r.entity_dofs[2]) if r.tdim == 3 else 0}
    for sub in SUBS[cell]:
        m = closure_size(r, sub)
        for nc in (1, 257, 5003):
            d0 = rng.integers(0, 1 << 30, (nc, m)).astype(np.int32)
            ent = rng.integers(0, 1000, nc).astype(np.uint32)  # the reference's tests draw entity_info < 1000
            ci = rng.integers(0, 1 << 30, nc).astype(np.uint32)
            # entity with the right type for the cell_info overload
            idx = next(i for i in range(nents[sub]) if sub == INTERVAL or r.subentity_types[2][i] == sub)
            for inverse in (False, True):
                want = r.permute_subentity_closure_batch(d0.copy(), ent, sub, None, inverse)
                got = d0.copy()
                e.permute_subentity_closure_batch(got, ent, sub, None, inverse)  # host pointers
                assert np.array_equal(got, want)
                gd, gi = torch.from_numpy(d0.copy()).cuda(), torch.from_numpy(ent.view(np.int32)).cuda()
                e.permute_subentity_closure_batch(gd, gi, sub, None, inverse)  # device pointers
                assert np.array_equal(gd.cpu().numpy(), want)
                want = r.permute_subentity_closure_batch(d0.copy(), ci, sub, idx, inverse)
                got = d0.copy()
                e.permute_subentity_closure_batch(got, ci, sub, idx, inverse)
                assert np.array_equal(got, want)
            # round trip, as test_permute_subentity_closure_inverse in the reference
            got = d0.copy()
            e.permute_subentity_closure_batch(got, ent, sub)
            e.permute_subentity_closure_batch(got, ent, sub, inverse=True)
            assert np.array_equal(got, d0)


@pytest.mark.gpu
def test_closure_single_cell_api_and_errors(ref):
    r, e = make_pair(ref, P, TET, 3, EQ)
    for info, want in REFERENCE_VECTORS[0][3].items():
        data = np.arange(10, dtype=np.int32)
        out = e.permute_subentity_closure(data, info, TRI)
        assert list(out) == want and list(data) == list(range(10))  # the reference wrapper returns a copy
        assert list(e.permute_subentity_closure_inv(out, info, TRI)) == list(range(10))
    # cell_info overload: face 2 reflected and rotated once -> bits 6..8 = 0b011
    assert list(e.permute_subentity_closure(np.arange(10, dtype=np.int32), 0b011 << 6, TRI, 2)) == REFERENCE_VECTORS[0][3][3]
    # vertices: nothing moves
    assert list(e.permute_subentity_closure(np.arange(1, dtype=np.int32), 5, POINT)) == [0]
    with pytest.raises(RuntimeError):
        e.permute_subentity_closure(np.arange(9, dtype=np.int32), 1, TRI)  # wrong closure size
    with pytest.raises(RuntimeError):
        e.permute_subentity_closure(np.arange(10, dtype=np.int32), 1, TET)  # not a sub-entity dimension
    rn, en = make_pair(ref, 3, TET, 2, 11)  # N1curl: transformations are not permutations
    with pytest.raises(RuntimeError, match="not permutations"):
        en.permute_subentity_closure(np.arange(3, dtype=np.int32), 1, INTERVAL)
