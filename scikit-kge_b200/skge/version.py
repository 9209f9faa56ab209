"""Version of the B200 build; the leading component tracks the reference's API level (0.1)."""
version_info = (0, 1, 'b200', 1)
__version__ = '%d.%d+%s.%d' % version_info
