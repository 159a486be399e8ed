"""Headless stand-in for matplotlib, just enough for the reference's front-end (taumain.py:9-10, :62-89):
`pyplot.subplots / plot / text / show` and `animation.FuncAnimation`.  `show()` drives the animation the way the
GUI event loop would -- init once, then update(frame) -- until the data thread has delivered the last frame and
the producer process has exited; what the callbacks saw is written as JSON to $SQ_MPL_STUB_REPORT.
Test infrastructure (tests/test_taumain_headless.py); never imported by the product."""
