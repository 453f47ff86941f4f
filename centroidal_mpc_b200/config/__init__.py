"""Problem definitions with the attribute names of /root/reference/config/conf_*.py,
without the pinocchio / example_robot_data / robot_properties_solo imports (robot constants
are recorded as numbers; see ``_robots.py``)."""
