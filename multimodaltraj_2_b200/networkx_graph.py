"""Mirror of ``networkx_graph.py`` (reference :20-208): ``online_graph(args).ConstructGraph(...)``
returning a graph object whose ``get_node_attr('node_pos_list')`` yields per-pedestrian ``[8,2]``
position arrays, as train.py:74-87 consumes them.  NetworkX itself is not needed: the graph is a
dict of node attribute dicts.  The N x N ``dist_mat`` the reference allocates as ``zeros(1,1)`` and
never fills (:71) is produced for real by the pairwise CUDA kernel (``Graph.pairwise``)."""
from __future__ import annotations

import numpy as np


class Node():
    def __init__(self, node_id, node_pos_list):
        self.id, self.pos = node_id, node_pos_list
        self.targets, self.seq, self.vel = [], [], 0

    def setTargets(self, seq):
        self.targets = seq


class Graph():
    def __init__(self):
        self.nodes = {}
        self.edges = {}
        self.adj_mat, self.dist_mat = [], []
        self.Stateful, self.step = True, 0

    def getNodes(self):
        return self.nodes

    def getEdges(self):
        return self.edges

    def setNodes(self, framenum, node, pos_list_len=8):
        """networkx_graph.py:114-129: a known pedestrian gets row ``framenum`` written (rows >= 8 ignored);
        a first sighting only allocates a zero [8,2] array (reference defect F-5, kept)."""
        if node.id in self.nodes:
            if 0 <= framenum < len(self.nodes[node.id]['node_pos_list']):
                self.nodes[node.id]['node_pos_list'][framenum] = node.pos
        else:
            self.nodes[node.id] = dict(seq=node.seq, node_pos_list=np.zeros((pos_list_len, 2)), targets=node.targets,
                                       vel=node.vel)

    def get_node_attr(self, param):
        return {nid: attrs[param] for nid, attrs in self.nodes.items() if param in attrs}

    def pairwise(self, frame_row, r2=4.0, inv_2sigma2=0.5, device="cuda"):
        """The N x N pedestrian-distance kernel + adjacency of the nodes at ``frame_row`` (CUDA)."""
        import torch
        from . import ops
        pos = np.stack([v[frame_row] for v in self.get_node_attr('node_pos_list').values()]).astype(np.float32)
        n = len(pos)
        npad = (n + 3) // 4 * 4
        p = torch.zeros((1, npad, 2), dtype=torch.float32, device=device)
        p[0, :n] = torch.from_numpy(pos).to(device)
        valid = torch.zeros((1, npad), dtype=torch.uint8, device=device)
        valid[0, :n] = 1
        kern, adj, deg = ops.pairwise_adj(p, valid, r2, inv_2sigma2)
        self.dist_mat, self.adj_mat = kern[0, :n, :n], adj[0, :n, :n]
        return self.dist_mat, self.adj_mat, deg[0, :n]


class online_graph():
    def __init__(self, args):
        self.diff = args.obs_len
        self.nodes, self.edges = [{}], [{}]
        self.onlineGraph = Graph()

    def ConstructGraph(self, current_batch, future_traj, framenum, stateful=True, valid=False):
        """networkx_graph.py:30-73 (training branch): walk the batch's frames in order; the itr-th frame
        writes row ``itr`` of each of its pedestrians' position arrays."""
        g = self.onlineGraph
        g.step = framenum
        framenum = int(framenum)                 # the loader's frame pointer is a float (load_traj.py:141-142)
        self.pos_list_len = len(current_batch)
        for itr, key in enumerate(current_batch):
            for item in current_batch[key]:
                (ped, pos), = item.items()
                ped = int(ped)
                node = Node(ped, pos)
                if ped not in future_traj:
                    continue                                     # reference: KeyError -> continue (:67-68)
                tgt = future_traj[ped]
                if len(tgt) < framenum:
                    node.setTargets(tgt[0:12])
                elif len(tgt[framenum:framenum + 12]) < 12:
                    node.setTargets(tgt)
                else:
                    node.setTargets(tgt[framenum:framenum + 12])
                g.setNodes(itr, node)
        return g
