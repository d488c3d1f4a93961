def parallel_wrapper_fn(env_fn):
    """The golden generator drives CookingEnvironment below this wrapper."""

    def par_fn(**kwargs):
        return env_fn(**kwargs)

    return par_fn
