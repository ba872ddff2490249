def register(id, entry_point=None, **kwargs):
    import gymnasium

    gymnasium._REGISTRY[id] = entry_point
