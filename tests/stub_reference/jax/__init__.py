"""Stand-in for jax (tests/test_capture_tool.py only): the capture tool needs jax.random.PRNGKey."""


class random:  # noqa: N801
    @staticmethod
    def PRNGKey(seed):
        return int(seed)
