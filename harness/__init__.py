"""Measurement / test harness around the hot path (not product code): runs the reference's own, unmodified
consumers (yag_slam.graph_slam.GraphSlam) on top of a chosen `karto_scanmatcher` stand-in."""
