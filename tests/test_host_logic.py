"""Host-side bookkeeping of the C ABI on a staging-only handle (no GPU): numbering, duplicate
checks, prefix semantics and the reference's error texts
(methods_for_node_data_handle.rs:16-64, methods_for_truss_data_handle.rs:11-62 and siblings)."""
import numpy as np
import pytest

from finite_element_method_b200 import FEM, PLATE, FemError, meshes


def staged(n=16):
    return FEM(1e-4, 1e-12, n, device=-1)


def raises(code, text, fn, *a, **k):
    with pytest.raises(FemError) as e:
        fn(*a, **k)
    assert e.value.code == code, (e.value.code, str(e.value))
    assert str(e.value) == text, str(e.value)


def test_node_errors_match_reference_texts():
    f = staged(3)
    f.add_node(1, 0.0, 0.0, 0.0)
    f.add_node(2, 30.0, 0.0, 0.0)
    raises(1, "Node with number 1 already exists!", f.add_node, 1, 5.0, 0.0, 0.0)
    raises(4, "Node with coordinates x: 30.0, y: 0.0, z: 0.0 already exists!", f.add_node, 3, 30.0, 0.0, 0.0)
    raises(4, "Node with coordinates x: -0.0, y: 0.0, z: 0.0 already exists!", f.add_node, 3, -0.0, 0.0, 0.0)  # -0.0 == 0.0
    f.add_node(7, 1.5, -2e-7, 1e20)
    raises(5, "Nodes number could not be greater than 3!", f.add_node, 9, 1.0, 1.0, 1.0)
    assert f.counts() == (3, 0, 0, 0)


def test_rust_debug_float_formatting():
    f = staged(8)
    for i, (v, s) in enumerate([(0.1, "0.1"), (1e16, "1e16"), (123456.0, "123456.0"), (1.5e-5, "1.5e-5"),
                                (1e-4, "0.0001"), (-2.5, "-2.5"), (1 / 3, "0.3333333333333333")]):
        f.add_node(10 + i, v, float(i), 0.0)
        raises(4, f"Node with coordinates x: {s}, y: {float(i)}, z: 0.0 already exists!", f.add_node, 99, v, float(i), 0.0)


def test_truss_errors_and_order_of_checks():
    f = staged()
    f.add_nodes([1, 2, 3], [0, 1, 2], [0, 0, 0], [0, 0, 0])
    f.add_trusses([1], [1], [2], [1e6], [2.0])
    raises(2, "Node with number 9 does not exist!", f.add_trusses, [1], [9], [8], [1e6], [2.0])      # node 1 first
    raises(2, "Node with number 8 does not exist!", f.add_trusses, [1], [1], [8], [1e6], [2.0])
    raises(10, "Truss element with number 1 already exists!", f.add_trusses, [1], [2], [3], [1e6], [2.0])
    raises(11, "Truss element with node number 2 and 1 already exists!", f.add_trusses, [2], [2], [1], [1e6], [2.0])
    raises(20, "Young's modulus -1.0 is less or equal to zero!", f.add_trusses, [2], [2], [3], [-1.0], [2.0])
    raises(22, "Area 0.0 is less or equal to zero!", f.add_trusses, [2], [2], [3], [1.0], [0.0])
    raises(23, "Area2 -0.5 is less or equal to zero!", f.add_trusses, [2], [2], [3], [1.0], [1.0], [-0.5])
    # a failed add leaves nothing behind: the same number / pair can be used afterwards
    f.add_trusses([2], [2], [3], [1.0], [1.0], [np.nan])
    assert f.counts() == (3, 2, 0, 0)


def test_beam_and_plate_errors_keep_reference_quirks():
    f = staged()
    f.add_nodes([1, 2, 3, 4, 5], [0, 1, 1, 0, 5], [0, 0, 1, 1, 5], [0, 0, 0, 0, 0])
    ax = np.array([[0.0], [0.0], [1.0]])
    ok = dict(young_modulus=[2e11], poisson_ratio=[0.3], area=[1e-2], i11=[8e-6], i22=[4e-6], i12=[0.0],
              it=[1e-5], shear_factor=[5 / 6], local_axis_1=ax)
    f.add_beams([1], [1], [2], **ok)
    raises(10, "Beam element with number 1 already exists!", f.add_beams, [1], [2], [3], **ok)
    raises(11, "Beam element with node number 2 and 1 already exists!", f.add_beams, [2], [2], [1], **ok)
    # beam.rs:86-87 reports young_modulus inside the Poisson message
    raises(21, "Poisson's ratio 200000000000.0 is less or equal to zero!", f.add_beams, [2], [2], [3],
           **{**ok, "poisson_ratio": [0.0]})
    raises(24, "I11 -1.0 is less or equal to zero!", f.add_beams, [2], [2], [3], **{**ok, "i11": [-1.0]})
    raises(26, "It 0.0 is less or equal to zero!", f.add_beams, [2], [2], [3], **{**ok, "it": [0.0]})
    f.add_plates([1], [3], [4], [1], [2], [2e11], [0.3], [0.01], [5 / 6])
    raises(10, "Plate element with number 1 already exists!", f.add_plates, [1], [3], [4], [1], [5], [2e11], [0.3], [0.01], [5 / 6])
    raises(11, "Plate element with nodes numbers [1, 2, 3, 4] already exists!", f.add_plates, [2], [1], [2], [3], [4],
           [2e11], [0.3], [0.01], [5 / 6])
    # plate.rs:83-84 reports young_modulus inside the Thickness message
    raises(29, "Thickness 200000000000.0 is less or equal to zero!", f.add_plates, [2], [3], [4], [1], [5],
           [2e11], [0.3], [0.0], [5 / 6])
    assert f.counts() == (5, 0, 1, 1)


def test_batch_is_prefix_atomic():
    f = staged()
    f.add_nodes(np.arange(1, 7), np.arange(6.0), np.zeros(6), np.zeros(6))
    raises(11, "Truss element with node number 2 and 1 already exists!", f.add_trusses,
           [1, 2, 3, 4], [1, 2, 2, 3], [2, 3, 1, 4], [1.0] * 4, [1.0] * 4)
    assert f.counts() == (6, 2, 0, 0)          # elements 1 and 2 kept, 3 (duplicate) and 4 dropped
    f.add_trusses([3, 4], [3, 4], [4, 5], [1.0, 1.0], [1.0, 1.0])
    assert f.counts() == (6, 4, 0, 0)


def test_reset_drops_everything():
    f = staged(2)
    f.add_nodes([1, 2], [0, 1], [0, 0], [0, 0])
    f.add_trusses([1], [1], [2], [1.0], [1.0])
    f.reset(4)
    assert f.counts() == (0, 0, 0, 0)
    f.add_nodes([1, 2, 3, 4], [0, 1, 2, 3], [0, 0, 0, 0], [0, 0, 0, 0])
    f.add_trusses([1], [1], [2], [1.0], [1.0])
    assert f.counts() == (4, 1, 0, 0)


def test_compute_calls_fail_loudly_without_device():
    f = staged(2)
    f.add_nodes([1, 2], [0, 1], [0, 0], [0, 0])
    f.add_trusses([1], [1], [2], [1.0], [1.0])
    for fn in (f.symbolic, f.numeric, f.assemble, f.validate, lambda: f.get_truss_rotation_matrix_elements(1)):
        with pytest.raises(FemError) as e:
            fn()
        assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_large_batches_take_the_parallel_path_with_identical_semantics():
    """>= 32768 keys are indexed by all host cores; the first failing element must still be the one a
    sequential loop over the reference would hit first."""
    n = 60000
    f = staged(n + 5)
    x = np.arange(n, dtype=np.float64)
    f.add_nodes(np.arange(1, n + 1), x, np.zeros(n), np.zeros(n))
    assert f.counts()[0] == n
    # duplicate coordinates in the middle of a big node batch: prefix kept
    f2 = staged(n + 5)
    x2 = x.copy(); x2[41234] = x2[17]
    raises(4, "Node with coordinates x: 17.0, y: 0.0, z: 0.0 already exists!", f2.add_nodes,
           np.arange(1, n + 1), x2, np.zeros(n), np.zeros(n))
    assert f2.counts()[0] == 41234
    # chain of trusses i -> i+1; element 50000 repeats the node pair of element 123 (reversed)
    a = np.arange(1, n, dtype=np.uint32); b = a + 1
    a2, b2 = a.copy(), b.copy()
    a2[50000], b2[50000] = b[123], a[123]
    ne = len(a)
    raises(11, f"Truss element with node number {b[123]} and {a[123]} already exists!", f.add_trusses,
           np.arange(1, ne + 1), a2, b2, np.ones(ne), np.ones(ne))
    assert f.counts()[1] == 50000
    # the rejected tail left no trace: the same numbers / pairs can be added now
    f.add_trusses(np.arange(50001, ne + 1), a[50000:], b[50000:], np.ones(ne - 50000), np.ones(ne - 50000))
    assert f.counts()[1] == ne
    # several failures in one batch: the earliest element wins, and for one element the reference's
    # check order wins (number before node set before property)
    f3 = staged(n + 5)
    f3.add_nodes(np.arange(1, n + 1), x, np.zeros(n), np.zeros(n))
    num = np.arange(1, ne + 1); num[40000] = 7            # duplicate number at 40000
    E = np.ones(ne); E[40000] = -1.0; E[45000] = -1.0      # property failures at 40000 and 45000
    raises(10, "Truss element with number 7 already exists!", f3.add_trusses, num, a, b, E, np.ones(ne))
    assert f3.counts()[1] == 40000


def test_degenerate_plate_uses_subset_rule():
    f = staged()
    f.add_nodes([1, 2, 3, 4], [0, 1, 1, 0], [0, 0, 1, 1], [0, 0, 0, 0])
    f.add_plates([1], [3], [4], [1], [2], [2e11], [0.3], [0.01], [5 / 6])
    # plate.rs:1114-1118: every node of the new element belongs to plate 1 -> "same nodes"
    raises(11, "Plate element with nodes numbers [1, 1, 2, 3] already exists!", f.add_plates, [2], [1], [1], [2], [3],
           [2e11], [0.3], [0.01], [5 / 6])


def test_row_strip_generators_match_local_part():
    """meshes.*(rows=...) — what bench.py builds per rank — is exactly local_part() of the whole mesh."""
    import numpy as np
    from finite_element_method_b200 import meshes
    for world in (2, 3, 8):
        full = meshes.mixed_structure(24, 19)
        for b, e in meshes.partition_rows(full, world, 25):
            a = meshes.local_part(full, b, e)
            c = meshes.mixed_structure(24, 19, rows=(b // 25, e // 25))
            for k in ("x", "p_n", "p_props", "b_n1", "b_n2", "b_props", "b_axis", "t_n1", "t_n2", "t_E", "t_A"):
                assert np.array_equal(np.asarray(a[k]), np.asarray(c[k])), (world, b, e, k)
        full = meshes.plate_grid(17, 11, "flat")
        for b, e in meshes.partition_rows(full, world, 18):
            a = meshes.local_part(full, b, e)
            c = meshes.plate_grid(17, 11, "flat", rows=(b // 18, e // 18))
            for k in ("p_n", "p_props"):
                assert np.array_equal(np.asarray(a[k]), np.asarray(c[k])), (world, b, e, k)


def test_reset_reuses_storage_but_forgets_everything():
    """FEM::reset (fem.rs:155-169) on a re-used instance keeps the staging vectors and hash tables allocated but must
    forget every node, element, number and node set: the same model loads again, twice, and duplicates are still
    caught afterwards (both through the per-call path and the batched, multi-threaded one)."""
    mesh = meshes.mixed_structure(60, 40)             # > 32768 keys per batch: the sharded multi-thread inserts
    n = len(mesh["x"])
    fem = FEM(mesh["rel_tol"], mesh["abs_tol"], n + 2, device=-1)   # room for the two probes (the limit is checked first)
    for _ in range(3):
        fem.load_mesh(mesh)
        assert fem.counts() == (n, len(mesh["t_n1"]), len(mesh["b_n1"]), len(mesh["p_n"][0]))
        assert list(fem.node_numbers()[:3]) == [1, 2, 3] and len(fem.element_numbers(PLATE)) == len(mesh["p_n"][0])
        with pytest.raises(FemError, match="Node with number 1 already exists!"):
            fem.add_node(1, -5.0, -5.0, -5.0)
        with pytest.raises(FemError, match="already exists!"):
            fem.add_node(10 ** 6, float(mesh["x"][7]), float(mesh["y"][7]), float(mesh["z"][7]))   # same coordinates
        with pytest.raises(FemError, match="Plate element with nodes numbers"):
            pn = np.asarray(mesh["p_n"]).reshape(4, -1)[:, 5] + 1
            fem.add_plate(10 ** 6, int(pn[2]), int(pn[3]), int(pn[0]), int(pn[1]), 2.1e11, 0.3, 0.01, 5 / 6)      # same node set, rotated
        fem.reset(n + 2)
        assert fem.counts() == (0, 0, 0, 0)
    # a smaller model after a larger one: nothing of the old one may be found
    fem.reset(2)
    fem.add_node(1, 0.0, 0.0, 0.0)
    fem.add_node(2, 30.0, 0.0, 0.0)
    fem.add_trusses([1], [1], [2], [1e6], [2.0])         # the batched form: no device validation on a staging-only handle
    assert fem.counts() == (2, 1, 0, 0)
    with pytest.raises(FemError, match="Node with number 3 does not exist!"):
        fem.add_trusses([2], [1], [3], [1e6], [2.0])
    fem.close()


def test_labels_added_out_of_order_are_all_found():
    """a label inserted while the dense number table was still short went to the
    sparse map; once the dense table grew past it, lookups missed it. Label 100000 before 1..70000."""
    n = 70001
    f = staged(n + 8)
    f.add_node(100000, -1.0, 0.0, 0.0)
    k = np.arange(1, n, dtype=np.uint32)
    f.add_nodes(k, k.astype(np.float64), np.zeros(n - 1), np.zeros(n - 1))
    # the early label is still known: as a duplicate ...
    raises(1, "Node with number 100000 already exists!", f.add_node, 100000, -2.0, 0.0, 0.0)
    # ... and as an element node
    f.add_trusses([1], [100000], [1], [1e6], [2.0])
    f.add_trusses([2], [70000], [100000], [1e6], [2.0])
    assert f.counts() == (n, 2, 0, 0)
    # element labels go through the same map
    g = staged(16)
    g.add_nodes(np.arange(1, 11), np.arange(10.0), np.zeros(10), np.zeros(10))
    g.add_trusses([500000], [1], [2], [1e6], [2.0])
    lab = np.arange(1, 9, dtype=np.uint32)
    g.add_trusses(lab, lab + 1, lab + 2, np.full(8, 1e6), np.full(8, 2.0))
    raises(10, "Truss element with number 500000 already exists!", g.add_trusses, [500000], [9], [10], [1e6], [2.0])
    # a rolled-back batch forgets its labels again, wherever they lived
    h = staged(4)
    raises(5, "Nodes number could not be greater than 4!", h.add_nodes, [900000, 1, 2, 3, 4], np.arange(5.0),
           np.zeros(5), np.zeros(5))
    h2 = staged(8)
    h2.add_node(1, 0.0, 0.0, 0.0)
    # the batch fails at its FIRST node (coordinates exist): the labels 900000 and 5 it had already entered are forgotten
    raises(4, "Node with coordinates x: 0.0, y: 0.0, z: 0.0 already exists!", h2.add_nodes, [900000, 5], [0.0, 1.0],
           np.zeros(2), np.zeros(2))
    h2.add_nodes([5, 900000], [5.0, 6.0], np.zeros(2), np.zeros(2))
    raises(1, "Node with number 900000 already exists!", h2.add_node, 900000, 9.0, 0.0, 0.0)
    assert h2.counts()[0] == 3


def test_bulk_wrappers_reject_ragged_arrays():
    f = staged(8)
    with pytest.raises(ValueError):
        f.add_nodes([1, 2, 3], [0.0, 1.0], [0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
    f.add_nodes([1, 2, 3, 4], [0.0, 1.0, 1.0, 0.0], [0.0, 0.0, 1.0, 1.0], np.zeros(4))
    with pytest.raises(ValueError):
        f.add_trusses([1, 2], [1, 2], [2], [1e6, 1e6], [2.0, 2.0])
    with pytest.raises(ValueError):
        f.add_plates([1], [3], [4], [1], [2], [2.1e11, 1.0], [0.3], [0.01], [5 / 6])
    with pytest.raises(ValueError):
        f.add_beams([1, 2], [1, 2], [2, 3], [1.0, 1.0], [0.3], [1.0, 1.0], [1.0, 1.0], [1.0, 1.0], [0.0, 0.0],
                    [1.0, 1.0], [1.0, 1.0], np.zeros((3, 2)))
    assert f.counts() == (4, 0, 0, 0)


def test_node_window_numbers_nodes_from_its_base():
    """femgpu_dist_set_node_window: a rank is handed the nodes [base, base + n) of the whole model only; their global
    insertion indices (hence matrix rows) start at base, the node limit still counts the whole model."""
    m = meshes.mixed_structure(6, 5)                     # 7 x 6 = 42 nodes
    n = len(m["x"])
    begin, end = meshes.partition_rows(m, 3, 7)[1]       # the middle strip
    part = meshes.with_node_window(meshes.local_part(m, begin, end), begin, end)
    assert part["node_window_begin"] == begin and len(part["x"]) == (end - begin) + 7 and part["nodes_number"] == n
    f = staged(n)
    f.load_mesh(part)
    assert f.counts() == (len(part["x"]), len(part["t_n1"]), len(part["b_n1"]), part["p_n"].shape[1])
    # a node outside the window is unknown to this rank
    raises(2, "Node with number 1 does not exist!", f.add_trusses, [9999], [1], [begin + 1], [1e6], [2.0])
    # labels inside the window resolve to global indices: duplicate node pair of an existing truss is found
    t1, t2 = int(part["t_n1"][0]) + 1, int(part["t_n2"][0]) + 1
    raises(11, f"Truss element with node number {t1} and {t2} already exists!", f.add_trusses, [9998], [t1], [t2], [1e6], [2.0])
    # the limit is global: the window may be filled up to nodes_number, not beyond
    g = staged(10)
    g._check(g._L.femgpu_dist_set_node_window(g._h, 8))
    g.add_nodes([9, 10], [0.0, 1.0], [0.0, 0.0], [0.0, 0.0])
    raises(5, "Nodes number could not be greater than 10!", g.add_node, 11, 2.0, 0.0, 0.0)
    with pytest.raises(FemError):
        g.dist_set_node_window(3)                        # only before the first node
    g.reset(10)                                          # FEM::reset forgets the window
    g.add_nodes(np.arange(1, 11), np.arange(10.0), np.zeros(10), np.zeros(10))
    assert g.counts()[0] == 10


def test_parallel_label_claims_and_partitioned_indices_keep_sequential_semantics():
    """Batches of >= 65536 labels are claimed in the dense number table by all host cores (atomic min per label) and
    the coordinate / node-set indices are filled shard by shard after a partition of the batch: the outcome must be
    the one of the reference's one-by-one loop — the first failing position, its check order, a clean prefix."""
    n = 200_000
    rng = np.random.default_rng(5)
    lab = rng.permutation(np.arange(1, n + 1, dtype=np.uint32))       # labels in no order at all
    x = np.arange(n, dtype=np.float64)
    zero = np.zeros(n)
    f = staged(n + 8)
    f.add_nodes(lab, x, zero, zero)
    assert f.counts()[0] == n and np.array_equal(f.node_numbers(), lab)
    raises(1, f"Node with number {lab[77]} already exists!", f.add_node, int(lab[77]), -1.0, 0.0, 0.0)
    # one label three times in a batch, the later copies in other threads' chunks: the SECOND occurrence fails
    for first, second, third in ((10, 150_000, 199_999), (120_000, 120_001, 190_000), (0, 1, 2)):
        g = staged(n + 8)
        l2 = lab.copy(); l2[second] = l2[first]; l2[third] = l2[first]
        raises(1, f"Node with number {l2[first]} already exists!", g.add_nodes, l2, x, zero, zero)
        assert g.counts()[0] == second
        # nothing of the rejected tail stayed behind: its labels and coordinates are free
        g.add_nodes(lab[second:], x[second:], zero[second:], zero[second:])
        assert g.counts()[0] == n and np.array_equal(g.node_numbers(), lab)
    # a label of an EARLIER batch in the middle of a big one; a coordinate clash before it wins, after it loses
    g = staged(2 * n)
    g.add_nodes(np.arange(1, 1001), -1.0 - np.arange(1000.0), np.zeros(1000), np.zeros(1000))
    l3 = np.arange(1001, 1001 + n, dtype=np.uint32); l3[123_456] = 500
    x3 = x.copy(); x3[180_000] = x3[5]
    raises(1, "Node with number 500 already exists!", g.add_nodes, l3, x3, zero, zero)
    assert g.counts()[0] == 1000 + 123_456
    x3[100_000] = x3[5]
    g2 = staged(2 * n)
    g2.add_nodes(np.arange(1, 1001), -1.0 - np.arange(1000.0), np.zeros(1000), np.zeros(1000))
    raises(4, "Node with coordinates x: 5.0, y: 0.0, z: 0.0 already exists!", g2.add_nodes, l3, x3, zero, zero)
    assert g2.counts()[0] == 1000 + 100_000
    # both checks fail at ONE position: the number check comes first (methods_for_node_data_handle.rs:42-64)
    l4 = np.arange(1, n + 1, dtype=np.uint32); l4[150_000] = 9
    x4 = x.copy(); x4[150_000] = x4[3]
    g3 = staged(n + 8)
    raises(1, "Node with number 9 already exists!", g3.add_nodes, l4, x4, zero, zero)
    assert g3.counts()[0] == 150_000
    # strictly ascending labels above everything known take a path without claims; an ascending batch that reaches
    # back into known labels must still find its first clash
    g4 = staged(3 * n)
    g4.add_nodes(np.arange(1, n + 1), x, zero, zero)
    g4.add_nodes(np.arange(n + 1, 2 * n + 1), x + n, zero, zero)
    assert g4.counts()[0] == 2 * n
    raises(1, f"Node with number {2 * n - 4} already exists!", g4.add_nodes, np.arange(2 * n - 4, 3 * n - 4), x + 2 * n, zero, zero)
    l5 = np.arange(2 * n + 1, 3 * n + 1, dtype=np.uint32); l5[70_000] = l5[69_999]       # not ascending: a repeat
    raises(1, f"Node with number {l5[69_999]} already exists!", g4.add_nodes, l5, x + 2 * n, zero, zero)
    assert g4.counts()[0] == 2 * n + 70_000
    g4.reset(3 * n)                                                                    # the table is zeroed, not dropped
    g4.add_nodes(np.arange(5, n + 5), x, zero, zero)
    raises(1, "Node with number 5 already exists!", g4.add_node, 5, -1.0, 0.0, 0.0)
    g4.add_node(4, -1.0, 0.0, 0.0)
    assert g4.counts()[0] == n + 1
    # elements: plates on a strip, labels permuted; random duplicate positions against a sequential model
    m = 100_000
    h = staged(2 * m + 2)
    k = np.arange(m + 1)
    h.add_nodes(np.arange(1, 2 * m + 3), np.concatenate([k, k]).astype(np.float64), np.repeat([0.0, 1.0], m + 1), np.zeros(2 * m + 2))
    n1 = np.arange(1, m + 1, dtype=np.uint32); n2 = n1 + 1; n3 = n2 + (m + 1); n4 = n1 + (m + 1)
    props = [np.full(m, 2e11), np.full(m, 0.3), np.full(m, 0.01), np.full(m, 5 / 6)]
    for trial in range(4):
        el = rng.permutation(np.arange(1, m + 1, dtype=np.uint32))
        a, b, c, d = n1.copy(), n2.copy(), n3.copy(), n4.copy()
        pos = np.sort(rng.choice(np.arange(70_000, m), 3, replace=False))
        kind = rng.integers(0, 2, 3)
        for p, kd in zip(pos, kind):
            src = int(rng.integers(0, p))
            if kd == 0:
                el[p] = el[src]                                          # label of an earlier plate
            else:
                a[p], b[p], c[p], d[p] = n3[src], n1[src], n4[src], n2[src]   # its node set, permuted
        hh = staged(2 * m + 2)
        hh.add_nodes(np.arange(1, 2 * m + 3), np.concatenate([k, k]).astype(np.float64), np.repeat([0.0, 1.0], m + 1), np.zeros(2 * m + 2))
        with pytest.raises(FemError) as e:
            hh.add_plates(el, a, b, c, d, *props)
        assert e.value.code == (10 if kind[0] == 0 else 11), str(e.value)
        assert hh.counts()[3] == pos[0]
        assert np.array_equal(hh.element_numbers(PLATE), el[:pos[0]])


def test_label_arrays_of_a_mesh_are_cached_only_while_the_sources_stay_the_same_objects():
    """FEM.api_arrays(mesh, cache=True) (what bench.py's e2e steps pass to load_mesh): built once per mesh dict, rebuilt
    when a source array was replaced; without `cache` nothing is kept in the dict."""
    mesh = meshes.mixed_structure(6, 4)
    a = FEM.api_arrays(mesh)
    assert "_api" not in mesh
    b = FEM.api_arrays(mesh, cache=True)
    assert FEM.api_arrays(mesh, cache=True) is b and mesh["_api"] is b
    for x, y in zip(a["plates"], b["plates"]):
        assert np.array_equal(x, y)
    m2 = dict(mesh)                                   # a copy of the dict carries the cache along ...
    m2["p_props"] = mesh["p_props"] * 2.0             # ... but not past a replaced source array
    c = FEM.api_arrays(m2, cache=True)
    assert c is not b and np.array_equal(c["plates"][5], 2.0 * b["plates"][5])
    f, g = staged(len(mesh["x"])), staged(len(mesh["x"]))
    f.load_mesh(mesh, cache=True); g.load_mesh(mesh)
    assert f.counts() == g.counts() == (len(mesh["x"]), 12, 24, 24)
