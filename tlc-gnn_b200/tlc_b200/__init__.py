"""B200-native vicinity persistence diagrams + persistence images (the TLC-GNN sg2dgm hot path)."""
