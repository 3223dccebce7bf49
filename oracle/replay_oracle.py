"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference replay memory
(dqn_agent.py:8-65) in the (graph id, colours, action, reward, done) encoding the device ring uses.
Ring index = count % capacity (:33); a minibatch is the listed slots in order, each expanded to its
graph's adjacency bytes and colour vectors (what Batch.from_data_list yields for the complete
graphs, :55-56)."""
import numpy as np


class ReplayOracle(object):
    def __init__(self, capacity, adjs):
        self.cap = capacity
        self.adjs = [np.ascontiguousarray(a, dtype=np.uint8) for a in adjs]
        self.count = 0
        self.slots = [None] * capacity

    def store(self, graph_id, cur_colour, next_colour, action, reward, done):
        idx = self.count % self.cap
        self.slots[idx] = (int(graph_id), np.asarray(cur_colour, dtype=np.int32).copy(),
                           np.asarray(next_colour, dtype=np.int32).copy(), int(action), float(reward), int(done))
        self.count += 1
        return idx

    def gather(self, batch):
        t = [self.slots[int(i)] for i in batch]
        return dict(
            ns=np.asarray([self.adjs[g].shape[0] for g, *_ in t], dtype=np.int64),
            adj=np.concatenate([self.adjs[g].reshape(-1) for g, *_ in t]),
            cur_colour=np.concatenate([c for _, c, *_ in t]), next_colour=np.concatenate([x[2] for x in t]),
            action=np.asarray([x[3] for x in t], dtype=np.int32), reward=np.asarray([x[4] for x in t], dtype=np.float32),
            done=np.asarray([x[5] for x in t], dtype=np.int32))
