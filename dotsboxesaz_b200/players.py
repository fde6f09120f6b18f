"""Time-limited play process (reference: players.py:14-72).

Same protocol as the reference's `AZPlayer(mp.Process)`: requests `(game_uuid, game_state, generation, time_limit)` on
`requ_queue`, `None` as the poison pill; answers `(game_uuid, move)` on `resp_queue`, the move being the greedy argmax of
the visit counts of a search that runs until the time limit (`UCT_search(num_reads=1e12, ..., time_limit)`,
players.py:62-63), or `None` when the search visited nothing.

The search itself is the drop-in `mcts.UCT_search` on the CUDA engine; the net is called through the reference's
`AsyncBatchedProxy(NeuralNetWrapper)` seam, exactly as players.py:38-46 wires it.  `serve_once()` is the body of one
request, usable without a process (tests, embedding in another event loop).
"""
import asyncio

import torch.multiprocessing as mp


class AZPlayer(mp.Process):
    def __init__(self, params, time_limit, requ_queue, resp_queue):
        super().__init__()
        params.self_play.mcts.temperature = {0: 1e-5}  # players.py:17
        self.params = params
        self.models = {}
        self.time_limit = time_limit
        self.requ_queue = requ_queue
        self.resp_queue = resp_queue

    def _load_model(self, generation):
        """players.py:24-31: one model object per generation, loaded on first use."""
        if generation in self.models:
            return self.models[generation]
        model = self.params.nn.model_class(self.params)
        model.load_parameters(generation)
        self.models[generation] = model
        return model

    async def serve_once(self, nn, game_state, time_limit):
        """One request (players.py:60-70): search from `game_state` until `time_limit` seconds have passed, return the
        most visited legal move (numpy integer) or None."""
        from .mcts import UCT_search, create_root_uct_node
        mcts_cfg = self.params.self_play.mcts
        node = create_root_uct_node(game_state)
        policy = await UCT_search(node, int(1e12), nn, mcts_cfg.mcts_cpuct, mcts_cfg.max_async_searches, (0.0, 0.0),
                                  time_limit if time_limit is not None else self.time_limit)
        if policy.sum() > 0:
            policy = policy * game_state.get_valid_moves()
            return policy.argmax()
        return None

    def run(self):
        from .nn import NeuralNetWrapper
        from .utils.proxies import AsyncBatchedProxy
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
        params = self.params
        nn_wrapper = NeuralNetWrapper(None, params)
        nn = AsyncBatchedProxy(nn_wrapper, batch_size=params.self_play.nn_batch_size, timeout=params.self_play.nn_batch_timeout,
                               batch_builder=params.self_play.nn_batch_builder)
        nn_task = loop.create_task(nn.run())
        while True:
            req = self.requ_queue.get()
            if req is None:  # poison pill
                break
            game_uuid, game_state, generation, time_limit = req
            nn_wrapper.set_model(self._load_model(generation))
            move = loop.run_until_complete(self.serve_once(nn, game_state, time_limit))
            self.resp_queue.put((game_uuid, move))
        nn_task.cancel()
        loop.run_until_complete(asyncio.sleep(0.1))
        loop.close()
