"""CPU: what the directed known-answer traces (recorded from the unmodified reference by
`tests/golden/make_golden.py kat`) must show.  The replay tests (oracle, C oracle, CUDA) hold every implementation to
every array of these files; this file states, scenario by scenario, WHICH reference behaviour each step pins."""
import json
import os

import numpy as np

from tests.replay import GOLDEN_DIR

# canonical dynamic slots of example.json: Plate 0-2, Tomato 3-5, Onion 6-8, Lettuce 9-11, Carrot 12-14, Banana 15-17,
# Apple 18-20, Watermelon 21-23, Bread 24-27; objs columns: present, x, y, chopped, blend, free, cont_kind, cont_id, pos
PLATE1, BANANA, WATERMELON, BREAD0, BREAD2 = 1, 15, 21, 24, 26
HELD, STATIC, ON_PLATE = 0, 1, 2


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def test_appendix_e_coop_scenarios():
    g = _load("kat_coop_seed0")
    ag, ob, st, obs = g["agents"][0], g["objs"][0], g["statics"][0], g["obs"][0]
    assert int(g["length"][0]) == 20 and int(g["raised"][0]) == -1
    assert np.all(g["reward"][0] == -5 / 400) and not g["term"][0].any() and not g["trunc"][0].any()   # -0.0125 per step
    xy = lambda t, i: tuple(ag[t, i, :2])
    # E1: both target (1,3): both cancelled, yet both orientations changed (C-2, single-pass collision rule)
    assert xy(1, 0) == (1, 2) and xy(1, 1) == (1, 4) and (ag[1, 0, 2], ag[1, 1, 2]) == (3, 4)
    # E2: agents may swap cells (C-3)
    assert xy(2, 0) == (1, 3) and xy(2, 1) == (1, 2)
    # E3: a1 bumps the counter and stays; a0 targeted a1's cell and is cancelled
    assert xy(3, 0) == (1, 2) and xy(3, 1) == (1, 1)
    # E5: grab from a counter
    assert ag[4, 1, 3] == BANANA and ob[4, BANANA, 6] == HELD and tuple(ob[4, BANANA, 1:3]) == (5, 1)
    # E6: Blender accepts the fresh Banana: READY (bit 1), toggle off, Banana still fresh
    blender = 3 + 29       # static slots: Cutboard 0-2, Counter 3-31, Blender 32-33
    assert ag[5, 1, 3] == -1 and tuple(ob[5, BANANA, [1, 2, 4, 6]]) == (4, 6, 0, STATIC) and st[5, blender, 3] == 1
    # E7: execute: mashed within the same step (max_progress 0), toggle off again, NOT_USABLE
    assert ob[6, BANANA, 4] == 2 and st[6, blender, 3] == 0
    # E8: pick the mashed Banana up; its observation block relative to the holder: [0, 0, !done 0, chopped 0, mashed 1, 1]
    assert ag[7, 1, 3] == BANANA
    assert obs[7, 1, 180:186].tolist() == [0, 0, 0, 0, 1, 1]
    # E9: a counter that holds a Watermelon does not accept the Banana (merge branch 3 rejected)
    assert ag[8, 1, 3] == BANANA and np.array_equal(ob[8, WATERMELON], ob[7, WATERMELON])
    # E10: a Deliversquare accepts any object
    assert ag[9, 1, 3] == -1 and tuple(ob[9, BANANA, [1, 2, 6]]) == (4, 0, STATIC)
    # E11: ... and never releases it
    assert ag[10, 1, 3] == -1 and np.array_equal(ob[10], ob[9])
    # E12: holding an empty Plate: the mashed Banana is scooped onto it (merge branch 2), now at the agent's cell
    assert ag[12, 1, 3] == PLATE1 and tuple(ob[12, BANANA, [1, 2, 6, 7]]) == (4, 1, ON_PLATE, PLATE1)
    # E13: Plate[Banana] vs a fresh Watermelon: branch 2 is selected, the plate rejects fresh food, no fall-through
    assert ag[13, 1, 3] == PLATE1 and np.array_equal(ob[13, WATERMELON], ob[12, WATERMELON])
    # E14: plate onto the Deliversquare; CarrotBanana stays incomplete
    assert ag[14, 1, 3] == -1 and tuple(ob[14, PLATE1, [1, 2, 6]]) == (4, 0, STATIC) and not g["marks"][0, 14].any()
    # E17: chopping a Bread creates a second, chopped Bread on top of it; the next grab takes the new one
    cut1 = 1
    assert st[16, cut1, 3] == 1 and ob[16, BREAD0, 6] == STATIC
    assert tuple(ob[17, BREAD0, [0, 3, 5, 8]]) == (1, 1, 0, 0) and tuple(ob[17, BREAD2, [0, 3, 5, 8]]) == (1, 1, 1, 1)
    assert st[17, cut1, 3] == 0 and obs[17, 0, 228 + 10:228 + 15].tolist()[2:] == [0, 1, 1]     # third Bread slot filled
    assert ag[18, 0, 3] == BREAD2 and ob[18, BREAD0, 5] == 1
    assert ag[20, 0, 3] == BREAD0


def test_appendix_e_three_agents_and_switch():
    g = _load("kat_open4_three")      # E4: the third agent ends on the first one's cell
    assert g["agents"][0, 1, :, :2].tolist() == [[1, 2], [1, 4], [1, 2]]
    g = _load("kat_switch")           # E15 / C-4: standing still on a Switch toggles it every step
    lay = json.loads(str(g["layouts"]))[0]
    names = [f"{k}{i}" for k, v in lay["meta"] if k in ("Cutboard", "Counter", "Blender", "Deliversquare", "Block", "Switch")
             for i in range(v)]
    sw, blk = names.index("Switch0"), names.index("Block0")
    st = g["statics"][0]
    assert [int(st[t, sw, 3]) & 4 for t in range(6)] == [0, 4, 0, 4, 4, 0]       # switch_active
    assert [int(st[t, blk, 3]) & 8 for t in range(6)] == [0, 8, 0, 8, 8, 0]      # Block.walkable follows
    assert g["agents"][0, 1:, 0, :2].tolist() == [[4, 3], [4, 3], [4, 3], [3, 3], [4, 3]]


def test_appendix_e16_scripted_solve_rewards():
    """the reference's own cook solves TomatoLettuceSalad alone (max_steps 200): -5/200 per step, then
    recipe_reward + that penalty on the completing step, terminated and not truncated"""
    g = _load("heuristic_cfg1")
    for k, n in enumerate(g["length"]):
        assert np.all(g["reward"][k, :n - 1, 0] == -5 / 200) and g["reward"][k, n - 1, 0] == 20 + -5 / 200
        assert g["term"][k, n - 1, 0] == 1 and g["trunc"][k, n - 1, 0] == 0 and not g["term"][k, :n - 1].any()
