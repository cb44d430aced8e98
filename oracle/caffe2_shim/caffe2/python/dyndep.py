"""unused by the paths exercised here"""
