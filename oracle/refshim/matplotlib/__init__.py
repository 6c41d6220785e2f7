"""TEST INFRASTRUCTURE ONLY. Import stand-in for matplotlib so the unmodified reference can be imported
(only `pyplot.colormaps['jet']` is touched, and only by the light model: /root/reference/sucre/sucre.py:104)."""
