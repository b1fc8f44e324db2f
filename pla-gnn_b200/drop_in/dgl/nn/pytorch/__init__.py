from plagnn_b200.nn import SAGEConv  # noqa: F401
