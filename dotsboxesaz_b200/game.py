"""The game-state interface the search is written against (reference: game.py:1-74)."""


class GameState:
    __slots__ = []

    @staticmethod
    def init_static_fields(*args, **kwargs):
        """Set the static (per-process) game parameters, e.g. the board size."""

    def get_actions_size(self):
        raise NotImplementedError

    def get_valid_moves(self, as_indices=False):
        """Legal moves: bool mask over the actions, or a list of indices with as_indices=True."""
        raise NotImplementedError

    def get_result(self):
        """Score from the perspective of the player to move: 1 win, 0 draw, -1 loss, None if not over."""
        raise NotImplementedError

    def play_(self, move):
        """Play `move` in place; ValueError if it is illegal."""
        raise NotImplementedError

    def play(self, move):
        """Play `move` on a copy and return the copy."""
        raise NotImplementedError

    def get_features(self):
        """Input planes of the policy/value net for this position."""
        raise NotImplementedError

    def get_hash(self):
        """A value unique to the position, kept incrementally."""
        raise NotImplementedError

    def __hash__(self):
        return self.get_hash().__hash__()

    def __eq__(self, other):
        return self.get_hash() == other.get_hash()
