"""Import stub (test infrastructure): utils/common_utils.py imports decord.VideoReader for its data loading helpers only."""


class VideoReader:  # placeholder
    pass
