"""Boundary-primitive construction, CPU side: the oracle's literal std::map restatement of
Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834) + Shell/IMPLICIT_EULER.h:245-277 against the independent
vectorised numpy restatement the scene generator uses (codim_ipc_b200.scenes.surface_primitives / assemble)."""
import numpy as np
import pytest

from oracle import cipc_oracle as O


def _scenes():
    from codim_ipc_b200 import scenes
    return {
        "stack": lambda: scenes.cloth_stack(12, 3),
        "sphere": lambda: scenes.cloth_on_sphere(24, draped=True),
        "noodles": lambda: scenes.noodles(4, 30),
        "granules": lambda: scenes.granules(800, cloth_n=10),
        "mixed": lambda: scenes.mixed_small(),
    }


@pytest.mark.parametrize("name", list(_scenes()))
def test_literal_restatement_matches_scene_generator(name):
    sc = _scenes()[name]()
    b = O.build_boundary(sc["X"], sc["F"], rod=sc["rodE"], rodRadius=np.full(len(sc["rodE"]), sc["xi"]), particle=sc["particles"])
    assert np.array_equal(b["BN"], sc["BN"]) and np.array_equal(b["BE"], sc["BE"]) and np.array_equal(b["BT"], sc["BT"])
    assert tuple(b["codim"]) == tuple(sc["codim"])
    assert np.allclose(b["BTArea"], sc["BTArea"], rtol=1e-13, atol=0) and np.allclose(b["BEArea"], sc["BEArea"], rtol=1e-12, atol=1e-300)
    assert np.allclose(b["BNArea"], sc["BNArea"], rtol=1e-12, atol=1e-300)


def test_orientation_and_overwrite_rules():
    """two triangles that mention a shared edge in the SAME direction: the second mention overwrites the area (map[(a,b)] = ...);
    opposite directions add; the edge keeps the orientation of its first mention; edges come out in lexicographic order"""
    X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1]], float)
    F = np.array([[2, 0, 1], [1, 3, 2], [0, 1, 4]], np.int32)  # edge (0,1): first mention (0,1) in tri 0, again (0,1) in tri 2
    b = O.build_boundary(X, F)
    be = [tuple(e) for e in b["BE"]]
    assert be == sorted(be) and (0, 1) in be and (1, 0) not in be and (2, 0) in be and (1, 2) in be and (2, 1) not in be
    A = [0.5, 0.5, 0.5]
    i01, i12 = be.index((0, 1)), be.index((1, 2))
    assert b["BEArea"][i01] == A[2] / 3 / 2           # overwritten by the last same-direction mention
    assert b["BEArea"][i12] == (A[0] / 3 + A[1] / 3) / 2  # opposite directions add
    assert list(b["BN"]) == [0, 1, 2, 3, 4]


def test_seg_rod_particle_appends():
    X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0], [3, 0, 0], [4, 0, 0], [5, 5, 5], [6, 6, 6], [9, 9, 9]], float)
    F = np.array([[0, 1, 2]], np.int32)
    b = O.build_boundary(X, F, seg=[[6, 7]], rod=[[4, 3], [4, 5]], rodRadius=[0.1, 0.2], particle=[8])
    assert [tuple(e) for e in b["BE"][-3:]] == [(6, 7), (4, 3), (4, 5)]
    assert list(b["BN"]) == [0, 1, 2, 6, 7, 3, 4, 5, 8] and tuple(b["codim"]) == (5, 8)
    assert len(b["BNArea"]) == 3 + 3  # surface nodes + rod nodes; seg ends and particles carry no area entry
    a0, a1 = 1.0 * np.pi * 0.1 / 6, 1.0 * np.pi * 0.2 / 6
    assert len(b["BEArea"]) == 3 + 2  # seg edges carry no area entry either
    assert np.allclose(b["BEArea"][-2:], [a0 / 2, a1 / 2]) and np.allclose(b["BNArea"][-3:], [a0 / 2, a0 / 2 + a1 / 2, a1 / 2])
