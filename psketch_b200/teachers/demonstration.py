"""Drop-in mirror of teachers/base.py + teachers/demonstration.py, backed by the CUDA teacher.

    teacher = DemonstrationTeacher(config)
    action = teacher(task, state)                              demonstration.py:9-30
    goal_pos, action_seq = teacher.find_closest_resources(task, state)     base.py:27-34

``state`` is a psketch_b200.worlds.craft.CraftState; the action for the state's current task is
normally already cached by the batched flush that produced the state.
"""


class BaseTeacher(object):
    def __init__(self, config=None):
        self.config = config

    def find_incomplete_subtask(self, task, state):
        """teachers/base.py:10-25 (host recursion over cached ``satisfies`` results)."""
        if state.satisfies(task):
            return None
        if task.subtasks is None:
            return task
        for subtask in task.subtasks[:-1]:
            found = self.find_incomplete_subtask(subtask, state)
            if found is not None:
                return found
        found = self.find_incomplete_subtask(task.subtasks[-1], state)
        assert found is not None
        return found

    def find_closest_resources(self, task, state, seq_cap=96):
        """(goal_pos, action_seq) of the closest cell holding ``task.goal_arg``; (last goal cell,
        None) when none is reachable; (None, None) when the kind is absent."""
        world = state.world
        kind = world.cookbook.index[task.goal_arg] or 0
        goal, length, seq = world.backend().find_closest(state.cells, state._agent, kind, seq_cap)
        goal_pos = None if goal[0] == 255 else (int(goal[0]), int(goal[1]))
        if length < 0:
            return goal_pos, None
        return goal_pos, [int(a) for a in seq[:length]]

    def shortest_path(self, state, goal_pos):
        """teachers/base.py:36-87 for one explicit goal cell: restrict the search to that cell by
        hiding the other cells of its kind behind an inert marker."""
        world = state.world
        cells = state.cells.copy()
        w, h = world.WIDTH, world.HEIGHT
        gi = int(goal_pos[0]) * h + int(goal_pos[1])
        kind = int(cells[gi])
        if kind == 0:
            # the reference still searches for a state "facing" the empty cell; unsupported here
            raise ValueError("shortest_path to an empty cell is not supported")
        other = (cells == kind)
        other[gi] = False
        cells[other] = world.cookbook.index["boundary"]
        from ..worlds.craft import CraftState
        probe = CraftState(state.scenario, cells, state._agent.copy())
        task = type("T", (), {"goal_arg": world.cookbook.index.get(kind)})
        return self.find_closest_resources(task, probe)[1]


class DemonstrationTeacher(BaseTeacher):
    def __call__(self, task, state):
        return state.expert_action(task)
