"""Subset of the pycolab game engine, restated (test infrastructure only).

pycolab is a transitive, un-pinned dependency of the reference (it is reached
through ``gym.make`` at train.py:51 and ``env.step`` at
safe_grid_agents/common/learn.py:69) and is not on disk; this module restates
its published update/render algorithm as summarised in SURVEY.md section 8.1
("Engine step"):

  * a game is a backdrop plus *things* -- sprites (one cell) and drapes (a
    boolean curtain) -- keyed by one ASCII character each;
  * ``its_showtime()`` sets frame 0 and runs one update with ``actions=None``;
    ``play(a)`` increments the frame and runs one update with ``a``;
  * an update walks the *update groups* in schedule order; every thing in a
    group sees the board rendered after the previous group finished;
  * rendering paints the backdrop, then the things in z-order (later wins);
  * ``the_plot`` carries the frame number, the summed reward of the frame,
    the terminate flag and arbitrary user keys.
"""
import collections

import numpy as np

Position = collections.namedtuple("Position", ["row", "col"])
Observation = collections.namedtuple("Observation", ["board", "layers"])


class Plot(dict):
    """Per-game blackboard shared by all things (pycolab `the_plot`)."""

    def __init__(self):
        super().__init__()
        self.frame = 0
        self._reward = None
        self._discount = 1.0
        self._game_over = False

    def add_reward(self, reward):
        self._reward = reward if self._reward is None else self._reward + reward

    def terminate_episode(self, discount=0.0):
        self._game_over = True
        self._discount = discount

    def _take_reward(self):
        reward, self._reward = self._reward, None
        return reward


class Backdrop:
    def __init__(self, curtain, palette):
        self.curtain = curtain
        self.palette = palette

    def update(self, actions, board, layers, things, the_plot):
        pass


class Sprite:
    def __init__(self, corner, position, character):
        self.corner = corner
        self.position = position
        self.character = character
        self.visible = True

    def update(self, actions, board, layers, backdrop, things, the_plot):
        raise NotImplementedError


class Drape:
    def __init__(self, curtain, character):
        self.curtain = curtain
        self.character = character

    def update(self, actions, board, layers, backdrop, things, the_plot):
        raise NotImplementedError


class MazeWalker(Sprite):
    """A sprite that walks one cell at a time and is stopped by `impassable`
    characters on the currently rendered board (pycolab prefab MazeWalker)."""

    def __init__(self, corner, position, character, impassable, confined_to_board=True):
        super().__init__(corner, position, character)
        self.impassable = frozenset(ord(c) for c in impassable)
        self.confined_to_board = confined_to_board

    def _move(self, board, the_plot, drow, dcol):
        row, col = self.position.row + drow, self.position.col + dcol
        inside = 0 <= row < self.corner.row and 0 <= col < self.corner.col
        if not inside:
            if self.confined_to_board:
                return "edge"
        elif int(board[row, col]) in self.impassable:
            return chr(int(board[row, col]))
        self.position = Position(row, col)
        return None

    def _north(self, board, the_plot):
        return self._move(board, the_plot, -1, 0)

    def _south(self, board, the_plot):
        return self._move(board, the_plot, 1, 0)

    def _west(self, board, the_plot):
        return self._move(board, the_plot, 0, -1)

    def _east(self, board, the_plot):
        return self._move(board, the_plot, 0, 1)

    def _stay(self, board, the_plot):
        return None


class Engine:
    def __init__(self, rows, cols, backdrop, things, update_groups, z_order, all_chars):
        self.rows, self.cols = rows, cols
        self.backdrop = backdrop
        self.things = things
        self.update_groups = update_groups
        self.z_order = z_order
        self.all_chars = all_chars
        self.the_plot = Plot()
        self._board = None
        self._layers = None
        self.game_over = False
        self._showtime_done = False

    def _render(self):
        board = self.backdrop.curtain.copy()
        for ch in self.z_order:
            thing = self.things[ch]
            if isinstance(thing, Sprite):
                if thing.visible:
                    board[thing.position.row, thing.position.col] = ord(ch)
            else:
                board[thing.curtain] = ord(ch)
        self._board = board
        self._layers = {ch: board == ord(ch) for ch in self.all_chars}

    def _update_and_render(self, actions):
        self.backdrop.update(actions, self._board, self._layers, self.things, self.the_plot)
        for group in self.update_groups:
            for ch in group:
                self.things[ch].update(
                    actions, self._board, self._layers, self.backdrop, self.things, self.the_plot
                )
            self._render()
        reward = self.the_plot._take_reward()
        self.game_over = self.the_plot._game_over
        return Observation(self._board, self._layers), reward, self.the_plot._discount

    def its_showtime(self):
        assert not self._showtime_done
        self._showtime_done = True
        self.the_plot.frame = 0
        self._render()
        return self._update_and_render(None)

    def play(self, actions):
        assert self._showtime_done and not self.game_over
        self.the_plot.frame += 1
        return self._update_and_render(actions)


def ascii_art_to_game(art, what_lies_beneath, sprites=None, drapes=None,
                      backdrop=Backdrop, update_schedule=None, z_order=None):
    """Build an Engine from ASCII art (pycolab `ascii_art.ascii_art_to_game`).

    sprites / drapes map a character to ``factory(corner, position, char)`` /
    ``factory(curtain, char)``.  `update_schedule` is a flat list (one group) or
    a list of lists (several groups); default is sorted characters in one
    group.  `z_order` defaults to the flattened update schedule.
    """
    sprites = sprites or {}
    drapes = drapes or {}
    grid = np.array([[ord(c) for c in line] for line in art], dtype=np.uint8)
    rows, cols = grid.shape
    corner = Position(rows, cols)
    things = {}
    back = grid.copy()
    for ch, factory in sprites.items():
        where = np.argwhere(grid == ord(ch))
        assert len(where) == 1, "sprite %r must appear exactly once" % ch
        things[ch] = factory(corner, Position(int(where[0][0]), int(where[0][1])), ch)
        back[grid == ord(ch)] = ord(what_lies_beneath)
    for ch, factory in drapes.items():
        things[ch] = factory(grid == ord(ch), ch)
        back[grid == ord(ch)] = ord(what_lies_beneath)
    if update_schedule is None:
        update_schedule = sorted(things)
    if update_schedule and not isinstance(update_schedule[0], (list, tuple)):
        update_groups = [list(update_schedule)]
    else:
        update_groups = [list(g) for g in update_schedule]
    flat = [ch for g in update_groups for ch in g]
    assert sorted(flat) == sorted(things)
    if z_order is None:
        z_order = flat
    palette = sorted({chr(v) for v in np.unique(back)})
    all_chars = sorted(set(palette) | set(things))
    return Engine(rows, cols, backdrop(back, palette), things, update_groups, list(z_order), all_chars)
