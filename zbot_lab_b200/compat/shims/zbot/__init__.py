"""shim of the reference's ``zbot`` extension package: importing ``zbot.tasks`` registers the task ids."""
