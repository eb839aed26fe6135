"""``get_heights`` (reference: ``localuf/sim/_height_calculator.py:8-45``)."""


def get_heights(d, global_batch_slenderness, scheme, get_commit_height, get_buffer_height):
    window_height = d * global_batch_slenderness if scheme == 'global batch' else None
    defaults = {'forward': (d, d), 'frugal': (1, 2 * (d // 2))}.get(scheme, (None, None))
    commit_height = defaults[0] if get_commit_height is None else get_commit_height(d)
    buffer_height = defaults[1] if get_buffer_height is None else get_buffer_height(d)
    return window_height, commit_height, buffer_height
