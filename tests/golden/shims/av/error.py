class FFmpegError(Exception):
    pass


class EOFError(FFmpegError):
    pass


class InvalidDataError(FFmpegError):
    pass
