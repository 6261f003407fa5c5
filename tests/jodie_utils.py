"""A tiny dataset in the on-disk format the reference reads (TGN / JODIE preprocessing: `data/ml_<name>.csv` with
columns [index, u, i, ts, label, idx], `data/ml_<name>.npy` edge features with a zero row 0,
`data/ml_<name>_node.npy` node features; reference loader tiger/data/data_loader.py:316-404)."""
import os

import numpy as np
import pandas as pd

from www2023tiger_b200.synthetic import StreamShape, make_stream


def write_toy_dataset(root, name='toy', n_users=60, n_items=20, n_events=3000, efeat_dim=6, seed=3, node_feats=True):
    st = make_stream(StreamShape(name, n_users, n_items, n_events, efeat_dim, None, horizon=50000.), seed=seed)
    os.makedirs(os.path.join(root, 'data'), exist_ok=True)
    df = pd.DataFrame({'u': st.src, 'i': st.dst, 'ts': st.ts, 'label': st.labels.astype(np.float64), 'idx': st.eids})
    df.to_csv(os.path.join(root, 'data', f'ml_{name}.csv'))
    np.save(os.path.join(root, 'data', f'ml_{name}.npy'), st.efeats)
    if node_feats:                       # TGN writes an all-zero node table of the edge-feature width
        np.save(os.path.join(root, 'data', f'ml_{name}_node.npy'), np.zeros((st.n_nodes, efeat_dim), dtype=np.float32))
    return st
